"""GPU parity tests: the CUDA path, called through the C ABI (include/jinc_b200.h), against the CPU oracle
(oracle/jinc_oracle.c, itself pinned bit-exactly to the unmodified reference in test_oracle_vs_ref.py).

Bars (BASELINE.json north_star): window origins and phase indices EXACT; integer samples within +-1 LSB of the
reference's opt=0 path; float within 1e-5 (scale-relative).  Weight blocks are additionally required to be
bit-identical -- phase blocks and on-the-fly border weights alike.
"""
import numpy as np
import pytest

from common import FULL_CASES, SMALL_CASES, assert_plane_close, make_filter, make_planes, oracle_frame, oracle_tables

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi(native_built):
    from jinc_b200 import capi as c

    assert c.device_count() >= 1, "no CUDA device: the product has no CPU fallback"
    return c


@pytest.mark.parametrize("case", SMALL_CASES, ids=[c[0] for c in SMALL_CASES])
def test_table_matches_reference_table(capi, case):
    name, fmt, w, h, tw, th, kw = case
    planes = make_planes(fmt, w, h)
    _, otabs = oracle_frame(fmt, w, h, tw, th, planes, **kw)
    flt = make_filter(fmt, w, h, tw, th, **kw)
    assert flt.num_tables == len(otabs)
    rng = np.random.default_rng(1)
    for k, ot in enumerate(otabs):
        tv = flt.table(k)
        info = tv.info
        assert info.filter_size == ot.filter_size
        sx, phx, bx, _ = tv.axis(0)
        sy, phy, by, _ = tv.axis(1)
        # window origins: exact, for every output pixel (the table is separable, the reference's is per pixel)
        assert np.array_equal(ot.meta[..., 0], np.broadcast_to(sx[None, :], ot.meta.shape[:2]))
        assert np.array_equal(ot.meta[..., 1], np.broadcast_to(sy[:, None], ot.meta.shape[:2]))
        # border flags and quantised phase indices: exact
        assert np.array_equal(ot.border.astype(bool), by[:, None].astype(bool) | bx[None, :].astype(bool))
        assert np.array_equal(ot.phase[..., 0], np.broadcast_to(phx[None, :], ot.phase.shape[:2]))
        assert np.array_equal(ot.phase[..., 1], np.broadcast_to(phy[:, None], ot.phase.shape[:2]))
        # same partition of pixels into shared weight blocks as coeff_meta
        H, W = ot.meta.shape[:2]
        ys = rng.integers(0, H, 400)
        xs = rng.integers(0, W, 400)
        ids = np.array([tv.pixel_block(int(x), int(y)) for x, y in zip(xs, ys)])
        ref_ids = ot.meta[ys, xs, 2]
        for a in range(0, 400, 7):
            same_ref = ref_ids == ref_ids[a]
            same_dev = ids == ids[a]
            assert np.array_equal(same_ref, same_dev)
        # weights: bit-identical on a sample of interior and border pixels (corners, edges, centre)
        pts = [(0, 0), (W - 1, 0), (0, H - 1), (W - 1, H - 1), (W // 2, 0), (0, H // 2), (W // 2, H // 2),
               (W // 2 + 1, H // 2), (W // 2, H // 2 + 1), (W // 3, H - 2), (W - 2, H // 3)]
        pts += [(int(x), int(y)) for x, y in zip(xs[:12], ys[:12])]
        # border strips, where pixels are folded into classes of identical blocks: walk along every edge
        pts += [(int(x), y) for x in rng.integers(0, W, 10) for y in (0, 1, 2, H - 3, H - 2, H - 1)]
        pts += [(x, int(y)) for y in rng.integers(0, H, 10) for x in (0, 1, 2, W - 3, W - 2, W - 1)]
        for x, y in pts:
            got = tv.pixel_weights(x, y)
            ref = ot.block(y, x)
            assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(ref).view(np.uint32)), (name, k, x, y)
    flt.close()


@pytest.mark.parametrize("kind", ["noise", "gradient", "impulse", "constant"])
@pytest.mark.parametrize("case", SMALL_CASES, ids=[c[0] for c in SMALL_CASES])
def test_frame_matches_reference(capi, case, kind):
    name, fmt, w, h, tw, th, kw = case
    planes = make_planes(fmt, w, h, kind)
    ref, _ = oracle_frame(fmt, w, h, tw, th, planes, **kw)
    flt = make_filter(fmt, w, h, tw, th, **kw)
    got = flt.process(planes)
    for i, (g, r) in enumerate(zip(got, ref)):
        assert_plane_close(g, r, fmt.bits == 32, f"{name}/{kind}/plane{i}")
    assert flt.kernel_launches > 0
    flt.close()


def test_pitched_and_pinned_buffers(capi):
    """Host planes with padded pitches (AviSynth frames) and page-locked planes (direct DMA path) give the same result."""
    import torch

    name, fmt, w, h, tw, th, kw = SMALL_CASES[1]
    planes = make_planes(fmt, w, h)
    ref, _ = oracle_frame(fmt, w, h, tw, th, planes, **kw)
    flt = make_filter(fmt, w, h, tw, th, **kw)
    padded = []
    for p in planes:
        buf = np.zeros((p.shape[0], p.shape[1] + 37), p.dtype)
        buf[:, : p.shape[1]] = p
        padded.append(buf[:, : p.shape[1]])
    got = flt.process(padded)
    for g, r in zip(got, ref):
        assert_plane_close(g, r, False, "padded")
    pin_src = [torch.from_numpy(p.copy()).pin_memory() for p in planes]
    pin_dst = [torch.zeros(r.shape, dtype=torch.uint8).pin_memory() for r in ref]
    flt.process([t.numpy() for t in pin_src], [t.numpy() for t in pin_dst])
    for g, r in zip(pin_dst, ref):
        assert_plane_close(g.numpy(), r, False, "pinned")
    flt.close()


def test_submit_wait_pipeline_keeps_frames_apart(capi):
    name, fmt, w, h, tw, th, kw = SMALL_CASES[0]
    flt = make_filter(fmt, w, h, tw, th, **kw)
    frames = [make_planes(fmt, w, h, "noise", seed=s) for s in range(6)]
    outs = [flt.alloc_dst() for _ in frames]
    tickets = []
    for f, o in zip(frames, outs):
        if len(tickets) >= 3:  # default slots per device
            flt.wait(tickets.pop(0))
        tickets.append(flt.submit(f, o))
    for t in tickets:
        flt.wait(t)
    for f, o in zip(frames, outs):
        ref, _ = oracle_frame(fmt, w, h, tw, th, f, **kw)
        for g, r in zip(o, ref):
            assert_plane_close(g, r, False, "pipeline")
    flt.close()


EXPECTED_PATHS = {
    "c1_yv12_2x_tap3": 1, "c4_rgbps_2x_tap8": 1, "c5_420p10_quarter_tap6_blur": 2, "half_tap3_yv12": 2,
    "down2to3_tap4_y16": 3,
    # rational ratios with piecewise-periodic phases: the chunked-cells kernel (3:2, 4:3, 4x, and a pure shift at 1:1)
    "up1p5_tap3_420p8": 4, "up1p5_tap4_444p16_crop": 4, "up1p5_tap3_f32_y": 4, "up4to3_tap3_420p8": 4, "up4to3_tap4_y16": 4,
    "up4x_tap3_420p8": 4, "up4x_tap4_rgbp16": 4, "same_size_shift": 4, "y8_tap2_1p5x": 4,
    "up5to4_tap3_420p8": 4, "up9to4_tap4_y16": 4, "down3to4_tap3_444p10": 4, "up3x_tap3_f32_y": 4, "up5to3_tap4_rgbp8": 4,
    "up1p5_tap5_422p12": 4, "yv411_tap5": 4, "down2to3_tap3_420p8": 4, "down2to3_tap3_f32_y": 4,
    # no structure (or a window size the fast kernels are not instantiated for): the general kernel
    "irregular_up_crop_mpeg1": 0, "irregular_down_quant": 0, "f32_444_3x_tap16": 0, "down2to3_tap4_444p16_crop": 0,
}


def test_fast_path_selection(capi):
    """Which kernel family the table build plans for a geometry (jinc_table_info.fast_path)."""
    cases = {c[0]: c for c in SMALL_CASES}
    for name, want in EXPECTED_PATHS.items():
        _, fmt, w, h, tw, th, kw = cases[name]
        flt = make_filter(fmt, w, h, tw, th, **kw)
        got = flt.table(0).info.fast_path
        flt.close()
        assert got == want, f"{name}: fast path {got}, expected {want}"


def test_window_larger_than_plane_is_rejected(capi):
    from minihost import avs_host as ah

    with pytest.raises(capi.JincError, match="larger than"):
        make_filter(ah.Format("y", 8), 8, 8, 4, 4, tap=8)


@pytest.mark.parametrize("case", FULL_CASES, ids=[c[0] for c in FULL_CASES])
def test_full_size_properties(capi, case):
    """Every BASELINE config at its full, benchmarked size: constant in -> constant out (weights sum to 1, border
    classes and resident border weights included), and agreement with the oracle on row bands at the top, in the middle
    (tile seams) and at the bottom of every plane -- whole rows, so the left and right border columns are covered --
    plus a band cut out of the left and right border strips over the full height."""
    name, fmt, w, h, tw, th, kw = case
    is_float = fmt.bits == 32
    flt = make_filter(fmt, w, h, tw, th, **kw)
    shapes = flt.plane_shapes()
    consts = (0.25, -0.125, 0.4, 0.7) if is_float else tuple(int(v * fmt.peak) for v in (0.3, 0.5, 0.8, 1.0))
    const = [np.full(s, v, fmt.dtype) for (s, _), v in zip(shapes, consts)]
    out = flt.process(const)
    for i, (o, v) in enumerate(zip(out, consts)):
        if is_float:
            assert np.abs(o - np.float32(v)).max() <= 1e-5, f"{name}: constant plane {i}"
        else:
            assert o.min() == v and o.max() == v, f"{name}: constant plane {i}: {o.min()}..{o.max()} != {v}"
    planes = make_planes(fmt, w, h)
    got = flt.process(planes)
    tabs = oracle_tables(fmt, w, h, tw, th, **kw)
    peak = float(fmt.peak) if fmt.bits < 32 else 0.0
    for i, pl in enumerate(planes):
        t = tabs[1] if (len(tabs) > 1 and i in (1, 2)) else tabs[0]
        H, W = t.dst_h, t.dst_w
        info = flt.table(1 if (len(tabs) > 1 and i in (1, 2)) else 0).info
        bands = [(0, 5), (H // 2 - 3, H // 2 + 3), (H - 5, H)]
        bands += [(y, y + 2) for y in (63, 64, 255, 256, H // 3, 2 * H // 3) if y + 2 <= H]  # tile seams of the fast paths
        bands += [(max(0, info.interior_y0 - 2), info.interior_y0 + 2), (info.interior_y1 - 2, min(H, info.interior_y1 + 2))]
        for (y0, y1) in bands:
            ref = t.resize(pl, peak, rows=(y0, y1))
            assert_plane_close(got[i][y0:y1], ref[y0:y1], is_float, f"{name}/plane{i}/rows{y0}-{y1}")
        # the left and right border strips over the full height: every 37th row, compared in the strip columns only
        rows = list(range(0, H, 37))
        xl, xr = info.interior_x0 + 8, max(info.interior_x1 - 8, 0)
        for y in rows:
            ref = t.resize(pl, peak, rows=(y, y + 1))
            assert_plane_close(got[i][y:y + 1, :xl], ref[y:y + 1, :xl], is_float, f"{name}/plane{i}/left/row{y}")
            assert_plane_close(got[i][y:y + 1, xr:], ref[y:y + 1, xr:], is_float, f"{name}/plane{i}/right/row{y}")
    for t in tabs:
        t.close()
    flt.close()


@pytest.mark.parametrize("dtype,bits,tap", [(np.uint8, 8, 3), (np.uint16, 16, 4), (np.float32, 32, 3)])
def test_device_plane_call_with_unaligned_source(capi, dtype, bits, tap):
    """jinc_resize_plane_device on planes the caller owns: a source whose base and pitch are NOT multiples of four
    samples must take the scalar tile-staging path of the exact-2x kernel and still match the oracle (the frame
    pipeline's own buffers are always vector-aligned, so only this entry point reaches it)."""
    import torch

    from oracle import cpu as oc

    w, h, tw, th = 322, 181, 644, 362
    sb = np.dtype(dtype).itemsize
    fmt_peak = float((1 << bits) - 1) if bits < 32 else 0.0
    rng = np.random.default_rng(7)
    if bits == 32:
        plane = rng.random((h, w), dtype=np.float32)
    else:
        plane = rng.integers(0, (1 << bits), (h, w)).astype(dtype)
    pp = oc.plane_params(w, h, tw, th, tap=tap, sub_w=0, sub_h=0)
    ot = oc.Table(pp[0], oc.make_lut(tap, 0.0))
    ref = ot.resize(plane, fmt_peak)

    ctx = capi.Context(0)
    tab = capi.Table(ctx, src_w=w, src_h=h, dst_w=tw, dst_h=th, radius=oc.radius_for_tap(tap))
    assert tab.info.fast_path == 1
    tdt = {1: torch.uint8, 2: torch.uint16, 4: torch.float32}[sb]
    for offset, pitch_elems in ((0, 336), (1, 325), (3, 323)):  # aligned, then two unaligned layouts
        buf = torch.zeros(offset + pitch_elems * h + 8, dtype=tdt, device="cuda")
        view = buf[offset:offset + pitch_elems * h].view(h, pitch_elems)
        view[:, :w].copy_(torch.from_numpy(plane))
        dpitch = (tw * sb + 63) // 64 * 64
        dst = torch.zeros((th, dpitch // sb), dtype=tdt, device="cuda")
        tab.resize_device(sb, fmt_peak, buf.data_ptr() + offset * sb, pitch_elems * sb, dst.data_ptr(), dpitch)
        torch.cuda.synchronize()
        got = dst[:, :tw].cpu().numpy()
        assert_plane_close(got, ref, bits == 32, f"device plane offset={offset} pitch={pitch_elems}")
    tab.close()
