"""Shared helpers for the parity tests: seeded synthetic planes, tolerances, case lists."""
from __future__ import annotations

import numpy as np

from minihost import avs_host as ah


def make_planes(fmt: ah.Format, width: int, height: int, kind: str = "noise", seed: int = 0):
    """Seeded synthetic frame: noise / gradient / impulse / constant."""
    rng = np.random.default_rng(0x4A494E43 ^ seed)
    planes = []
    for i in range(len(fmt.planes)):
        h, w = fmt.plane_shape(i, width, height)
        chroma = i in (1, 2) and fmt.family not in ("rgbp", "rgbap")
        if fmt.bits == 32:
            if kind == "noise":
                a = rng.random((h, w), dtype=np.float32) - (0.5 if chroma else 0.0)
            elif kind == "gradient":
                a = ((np.arange(w)[None, :] + np.arange(h)[:, None]) / float(w + h)).astype(np.float32)
            elif kind == "impulse":
                a = np.zeros((h, w), np.float32)
                a[::16, ::16] = 1.0
            else:
                a = np.full((h, w), 0.25, np.float32)
        else:
            peak = fmt.peak
            if kind == "noise":
                a = rng.integers(0, peak + 1, (h, w))
            elif kind == "gradient":
                a = ((np.arange(w)[None, :] + np.arange(h)[:, None]) * peak) // (w + h)
            elif kind == "impulse":
                a = np.zeros((h, w), np.int64)
                a[::16, ::16] = peak
            else:
                a = np.full((h, w), (peak + 1) // 2 if chroma else (peak + 1) // 16)
            a = a.astype(fmt.dtype)
        planes.append(np.ascontiguousarray(a))
    return planes


def assert_plane_close(got: np.ndarray, ref: np.ndarray, is_float: bool, what: str = ""):
    """The parity bar of BASELINE.json: integer output within +-1 LSB of the opt=0 reference, float within
    1e-5 relative to max(1, |ref|) (the reference's own SIMD paths miss element-wise 1e-5 near zero)."""
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    if is_float:
        err = np.abs(got.astype(np.float64) - ref.astype(np.float64))
        tol = 1e-5 * np.maximum(1.0, np.abs(ref.astype(np.float64)))
        bad = err > tol
        assert not bad.any(), f"{what}: {int(bad.sum())} samples beyond 1e-5 (max err {err.max():.3e})"
        return float(err.max())
    d = np.abs(got.astype(np.int64) - ref.astype(np.int64))
    assert d.max() <= 1, f"{what}: max |diff| = {int(d.max())} LSB at {np.unravel_index(d.argmax(), d.shape)}"
    return int((d != 0).sum())


# (name, format, src w,h, dst w,h, kwargs) -- reduced-size versions of the five BASELINE configs plus irregular cases
SMALL_CASES = [
    ("c1_yv12_2x_tap3", ah.YV12, 160, 90, 320, 180, dict(tap=3)),
    ("c2_420p8_2x_tap3_mpeg2", ah.YUV420P8, 480, 270, 960, 540, dict(tap=3, cplace="MPEG2")),
    ("c3_444p16_2x_tap4_crop", ah.YUV444P16, 240, 136, 480, 272, dict(tap=4, src_left=10.3, src_top=6.7, quant_x=256, quant_y=256)),
    ("c4_rgbps_2x_tap8", ah.RGBPS, 192, 108, 384, 216, dict(tap=8)),
    ("c5_420p10_quarter_tap6_blur", ah.YUV420P10, 768, 432, 192, 108, dict(tap=6, blur=0.9)),
    ("irregular_up_crop_mpeg1", ah.YV12, 320, 180, 500, 282, dict(tap=3, src_left=3.3, src_top=1.7, src_width=300.5, src_height=170.25, cplace="mpeg1")),
    ("irregular_down_quant", ah.Format("422", 10), 320, 180, 214, 120, dict(tap=4, quant_x=100, quant_y=37, blur=0.9)),
    ("topleft_negative_crop", ah.Format("420", 16), 320, 180, 212, 120, dict(tap=4, cplace="topleft", src_width=-10.5, src_height=-3.25)),
    ("y8_tap2_1p5x", ah.Format("y", 8), 200, 120, 300, 180, dict(tap=2)),
    ("yv411_tap5", ah.Format("411", 8), 320, 96, 480, 144, dict(tap=5)),
    ("rgbap12_2x_tap6", ah.Format("rgbap", 12), 128, 72, 256, 144, dict(tap=6)),
    ("yuva420_14bit_2x", ah.Format("yuva420", 14), 128, 72, 256, 144, dict(tap=3, cplace="topleft")),
    ("f32_444_3x_tap16", ah.Format("444", 32), 96, 64, 288, 192, dict(tap=16)),
    ("same_size_shift", ah.Format("y", 16), 160, 90, 160, 90, dict(tap=3, src_left=0.37, src_top=-0.21)),
    # integer-ratio downscales (polyphase kernel): every instantiated (ratio, filter size) and sample type
    ("half_tap3_yv12", ah.YV12, 640, 360, 320, 180, dict(tap=3)),
    ("half_tap4_f32_444", ah.Format("444", 32), 288, 160, 144, 80, dict(tap=4)),
    ("half_tap6_y16", ah.Format("y", 16), 400, 240, 200, 120, dict(tap=6)),
    ("half_tap8_rgbp10", ah.Format("rgbp", 10), 320, 200, 160, 100, dict(tap=8, blur=1.1)),
    ("quarter_tap3_y8", ah.Format("y", 8), 800, 480, 200, 120, dict(tap=3)),
    ("quarter_tap4_422p12", ah.Format("422", 12), 1024, 512, 256, 128, dict(tap=4)),
    ("quarter_tap6_f32_y", ah.Format("y", 32), 640, 400, 160, 100, dict(tap=6)),
    # exact 2x with the remaining alias taps and a third-size / 3x ratio (general kernel)
    ("up2x_tap6_420p8", ah.YUV420P8, 200, 120, 400, 240, dict(tap=6, cplace="mpeg1")),
    ("up2x_tap5_444p10", ah.Format("444", 10), 160, 96, 320, 192, dict(tap=5)),
    ("up2x_tap7_f32_y", ah.Format("y", 32), 160, 96, 320, 192, dict(tap=7, src_left=0.5)),
    ("third_tap3_y8", ah.Format("y", 8), 600, 360, 200, 120, dict(tap=3)),
    # 3:2 upscale (720p -> 1080p class): NOT periodic under the reference's float-accumulated positions (4+ phases per axis)
    ("up1p5_tap3_420p8", ah.YUV420P8, 320, 180, 480, 270, dict(tap=3)),
    ("up1p5_tap4_444p16_crop", ah.YUV444P16, 240, 136, 360, 204, dict(tap=4, src_left=1.25, src_top=0.75)),
    ("up1p5_tap3_f32_y", ah.Format("y", 32), 200, 120, 300, 180, dict(tap=3)),
    # 2:3 downscale (1080p -> 720p class): exactly periodic (source step 1.5): tap 3 on the chunked-cells kernel (one staging
    # per tile), tap 4 as four passes of the odd-ratio polyphase kernel
    ("down2to3_tap3_420p8", ah.YUV420P8, 480, 270, 320, 180, dict(tap=3)),
    ("down2to3_tap4_y16", ah.Format("y", 16), 300, 180, 200, 120, dict(tap=4)),
    ("down2to3_tap4_444p16_crop", ah.YUV444P16, 360, 204, 240, 136, dict(tap=4, src_left=1.5, src_top=0.75, src_width=357.0, src_height=202.5)),
    ("down2to3_tap3_f32_y", ah.Format("y", 32), 300, 180, 200, 120, dict(tap=3)),
    ("down2to3_tap3_422p10_mpeg1", ah.Format("422", 10), 384, 216, 256, 144, dict(tap=3, cplace="mpeg1")),
    # 4:3 upscale (1080p -> 1440p class): exactly periodic (source step 0.75), sixteen passes
    ("up4to3_tap3_420p8", ah.YUV420P8, 360, 204, 480, 272, dict(tap=3)),
    ("up4to3_tap4_y16", ah.Format("y", 16), 270, 150, 360, 200, dict(tap=4)),
    # 4x upscale: exactly periodic (source step 0.25), sixteen unit-step passes
    ("up4x_tap3_420p8", ah.YUV420P8, 160, 90, 640, 360, dict(tap=3)),
    ("up4x_tap4_rgbp16", ah.Format("rgbp", 16), 120, 68, 480, 272, dict(tap=4)),
    # more rational ratios of the chunked-cells kernel: 5:4 and 9:4 upscales and the 3:4 downscale (source step 4 per cell),
    # 3x (step 1), 5:3 (step 3), taps 2 and 5 at 3:2
    ("up5to4_tap3_420p8", ah.YUV420P8, 256, 144, 320, 180, dict(tap=3)),
    ("up9to4_tap4_y16", ah.Format("y", 16), 128, 64, 288, 144, dict(tap=4)),
    ("down3to4_tap3_444p10", ah.Format("444", 10), 320, 240, 240, 180, dict(tap=3)),
    ("up3x_tap3_f32_y", ah.Format("y", 32), 120, 72, 360, 216, dict(tap=3)),
    ("up5to3_tap4_rgbp8", ah.Format("rgbp", 8), 144, 96, 240, 160, dict(tap=4)),
    ("up1p5_tap5_422p12", ah.Format("422", 12), 256, 120, 384, 180, dict(tap=5, cplace="mpeg1")),
    # general kernel with four planes on one table, and a steep irregular downscale whose source footprints do not fit
    # in shared memory (per-plane fallback of the general kernel)
    ("rgbap10_irregular_up", ah.Format("rgbap", 10), 200, 120, 290, 170, dict(tap=4)),
    ("steep_down_444p16_tap5", ah.Format("444", 16), 900, 540, 200, 124, dict(tap=5)),
]


# The five BASELINE.json configs at their full sizes (config 1 is small already): (name, format, src, dst, kwargs)
FULL_CASES = [
    ("config1", ah.YV12, 640, 360, 1280, 720, dict(tap=3)),
    ("config2", ah.YUV420P8, 1920, 1080, 3840, 2160, dict(tap=3, cplace="MPEG2")),
    ("config3", ah.YUV444P16, 1920, 1080, 3840, 2160, dict(tap=4, src_left=10.3, src_top=6.7, quant_x=256, quant_y=256)),
    ("config4", ah.RGBPS, 3840, 2160, 7680, 4320, dict(tap=8)),
    ("config5", ah.YUV420P10, 7680, 4320, 1920, 1080, dict(tap=6, blur=0.9)),
    # the rational ratios of bench.py's workloads 6 and 9 at full size: the phase of a residue steps every few dozen cells
    ("720p_to_1080p", ah.YUV420P8, 1280, 720, 1920, 1080, dict(tap=3)),
    ("1080p_to_1440p_tap4", ah.Format("444", 10), 1920, 1080, 2560, 1440, dict(tap=4)),
]


def oracle_tables(fmt: ah.Format, w, h, tw, th, **kw):
    """The oracle's coefficient tables of a filter (one, or luma + chroma)."""
    from oracle import cpu as oc

    sw, sh = fmt.subsampling
    pp = oc.plane_params(w, h, tw, th, src_left=kw.get("src_left", 0.0), src_top=kw.get("src_top", 0.0),
                         src_width=kw.get("src_width"), src_height=kw.get("src_height"),
                         quant_x=kw.get("quant_x", 256), quant_y=kw.get("quant_y", 256), tap=kw.get("tap", 3),
                         sub_w=sw, sub_h=sh, cplace=kw.get("cplace", "mpeg2"))
    lut = oc.make_lut(kw.get("tap", 3), kw.get("blur", 0.0))
    return [oc.Table(p, lut) for p in pp]


def oracle_frame(fmt: ah.Format, w, h, tw, th, planes, **kw):
    """Reference result via the CPU oracle (oracle/jinc_oracle.c): returns (out_planes, tables)."""
    from oracle import cpu as oc

    sw, sh = fmt.subsampling
    pp = oc.plane_params(w, h, tw, th, src_left=kw.get("src_left", 0.0), src_top=kw.get("src_top", 0.0),
                         src_width=kw.get("src_width"), src_height=kw.get("src_height"),
                         quant_x=kw.get("quant_x", 256), quant_y=kw.get("quant_y", 256), tap=kw.get("tap", 3),
                         sub_w=sw, sub_h=sh, cplace=kw.get("cplace", "mpeg2"))
    lut = oc.make_lut(kw.get("tap", 3), kw.get("blur", 0.0))
    tables = [oc.Table(p, lut) for p in pp]
    outs = []
    for i, pl in enumerate(planes):
        t = tables[1] if (len(tables) > 1 and i in (1, 2)) else tables[0]
        outs.append(t.resize(pl, float(fmt.peak) if fmt.bits < 32 else 0.0))
    return outs, tables


def make_filter(fmt: ah.Format, w, h, tw, th, devices=(0,), **kw):
    from jinc_b200 import capi

    sw, sh = fmt.subsampling
    return capi.Filter(src_w=w, src_h=h, target_w=tw, target_h=th, n_planes=len(fmt.planes),
                       sample_bytes=np.dtype(fmt.dtype).itemsize, bits=fmt.bits, sub_w=sw, sub_h=sh,
                       src_left=kw.get("src_left", 0.0), src_top=kw.get("src_top", 0.0), src_width=kw.get("src_width"),
                       src_height=kw.get("src_height"), quant_x=kw.get("quant_x", 256), quant_y=kw.get("quant_y", 256),
                       tap=kw.get("tap", 3), blur=kw.get("blur", 0.0), cplace=kw.get("cplace", "mpeg2"),
                       devices=list(devices) if devices is not None else None, flags=kw.get("flags", 0),
                       slots_per_device=kw.get("slots_per_device", 0))
