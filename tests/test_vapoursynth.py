"""The VapourSynth (API 4) front-end over the same C ABI: registration, argument validation and
error texts without a GPU; frames against the CPU oracle and frame properties on the GPU.  The plugin is driven by the
in-process core stand-in (minihost/vs_minihost.cpp) through VapourSynth's own call sequence."""
import numpy as np
import pytest

from common import assert_plane_close, make_planes, oracle_frame
from minihost import avs_host as ah


@pytest.fixture(scope="module")
def vs(native_built):
    from jinc_b200 import paths
    from minihost import vs_host

    core = vs_host.Core()
    plugin = core.load_plugin(paths.vs_plugin())
    return vs_host, core, plugin


def _source(vs_host, core, fmt: ah.Format, w, h, frames, **kw):
    fam = {"y": vs_host.CF_GRAY, "rgbp": vs_host.CF_RGB}.get(fmt.family, vs_host.CF_YUV)
    ssw, ssh = fmt.subsampling
    if fmt.family == "rgbp":  # AviSynth order G,B,R -> VapourSynth order R,G,B
        frames = [[p[2], p[0], p[1]] for p in frames]
    return core.source(fam, fmt.dtype, fmt.bits, ssw, ssh, w, h, frames, **kw)


def test_vs_plugin_registers_jincresize(vs):
    _, _, plugin = vs
    assert plugin.namespace == "jinc"
    args = plugin.function_args("JincResize")
    assert args.startswith("clip:vnode;width:int;height:int;tap:int:opt;")
    for name in ("src_left", "src_top", "src_width", "src_height", "blur"):
        assert f"{name}:float:opt;" in args
    for name in ("quant_x", "quant_y"):
        assert f"{name}:int:opt;" in args
    assert "cplace:data:opt;" in args


VS_ERRORS = [
    (dict(tap=0), "JincResize: tap must be between 1..16."),
    (dict(tap=17), "JincResize: tap must be between 1..16."),
    (dict(quant_x=0), "JincResize: quant_x must be between 1..256."),
    (dict(quant_y=257), "JincResize: quant_y must be between 1..256."),
    (dict(cplace="center"), "JincResize: cplace must be MPEG2, MPEG1 or topleft."),
]


@pytest.mark.parametrize("kw,msg", VS_ERRORS)
def test_vs_error_texts_are_the_references(vs, kw, msg):
    """Argument validation runs before any GPU work and uses the reference's messages (src/JincResize.cpp:703-723)."""
    vs_host, core, plugin = vs
    src = _source(vs_host, core, ah.Format("y", 8), 64, 64, [[np.zeros((64, 64), np.uint8)]])
    with pytest.raises(vs_host.VsError) as e:
        plugin.invoke("JincResize", src, width=96, height=96, **kw)
    assert str(e.value) == msg
    src.release()


def test_vs_topleft_only_for_420_and_required_arguments(vs):
    vs_host, core, plugin = vs
    f422 = ah.Format("422", 8)
    src = _source(vs_host, core, f422, 64, 64, [make_planes(f422, 64, 64)])
    with pytest.raises(vs_host.VsError, match="topleft must be used only for 4:2:0"):
        plugin.invoke("JincResize", src, width=96, height=96, cplace="topleft")
    with pytest.raises(vs_host.VsError, match="argument height is required"):
        plugin.invoke("JincResize", src, width=96)
    with pytest.raises(vs_host.VsError, match="does not take argument"):
        plugin.invoke("JincResize", src, width=96, height=96, opt=0)
    bad = _source(vs_host, core, f422, 64, 64, [make_planes(f422, 64, 64)], props={"_ChromaLocation": 5})
    with pytest.raises(vs_host.VsError, match="invalid _ChromaLocation"):
        plugin.invoke("JincResize", bad, width=96, height=96)
    src.release()
    bad.release()
    assert vs_host.Core.live_objects()[1] == 0  # failed constructions release the clip


VS_CASES = [
    ("420p8_2x", ah.YUV420P8, 160, 90, 320, 180, dict(tap=3, cplace="MPEG2")),
    ("444p16_crop", ah.YUV444P16, 120, 68, 240, 136, dict(tap=4, src_left=10.3, src_top=6.7)),
    ("rgbps_tap8", ah.RGBPS, 96, 54, 192, 108, dict(tap=8)),
    ("420p10_quarter_blur", ah.YUV420P10, 384, 216, 96, 54, dict(tap=6, blur=0.9)),
    ("gray8_1p5x", ah.Format("y", 8), 200, 120, 300, 180, dict(tap=3)),
    ("422p10_mpeg1_3to2", ah.Format("422", 10), 128, 72, 192, 108, dict(tap=3, cplace="mpeg1")),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", VS_CASES, ids=[c[0] for c in VS_CASES])
def test_vs_frames_match_oracle(vs, case):
    vs_host, core, plugin = vs
    name, fmt, w, h, tw, th, kw = case
    frames = [make_planes(fmt, w, h, "noise", seed=s) for s in range(2)]
    src = _source(vs_host, core, fmt, w, h, frames)
    clip = plugin.invoke("JincResize", src, width=tw, height=th, **kw)
    assert (clip.info["width"], clip.info["height"]) == (tw, th)
    for n in range(2):
        got, props = clip.get_frame(n)
        if fmt.family == "rgbp":
            got = [got[1], got[2], got[0]]  # back to the oracle's G,B,R order
        ref, _ = oracle_frame(fmt, w, h, tw, th, frames[n], **kw)
        for i, (g, r) in enumerate(zip(got, ref)):
            assert_plane_close(g, r, fmt.bits == 32, f"vs/{name}/frame{n}/plane{i}")
        if fmt.family in ("420", "422"):
            assert props["_ChromaLocation"] == {"mpeg2": 0, "mpeg1": 1, "topleft": 2}[kw.get("cplace", "mpeg2").lower()]
        else:
            assert "_ChromaLocation" not in props
    clip.release()
    src.release()
    assert vs_host.Core.live_objects() == (0, 0)


@pytest.mark.gpu
def test_vs_cplace_from_frame_property_and_parallel_pull(vs):
    """cplace defaults from frame 0's _ChromaLocation (as the AviSynth side does, :725-742); the filter is fmParallel:
    six host threads pull frames through one instance."""
    vs_host, core, plugin = vs
    fmt, w, h = ah.YV12, 96, 64
    frames = [make_planes(fmt, w, h, "noise", seed=s) for s in range(3)]
    for loc, cp in ((0, "mpeg2"), (1, "mpeg1"), (2, "topleft")):
        src = _source(vs_host, core, fmt, w, h, frames, props={"_ChromaLocation": loc, "_Matrix": 6}, num_frames=48)
        clip = plugin.invoke("JincResize", src, width=192, height=128)
        assert clip.pull(0, 36, threads=6) > 0
        got, props = clip.get_frame(7)
        ref, _ = oracle_frame(fmt, w, h, 192, 128, frames[7 % 3], tap=3, cplace=cp)
        for g, r in zip(got, ref):
            assert_plane_close(g, r, False, f"vs/cplace-from-prop/{cp}")
        assert props == {"_ChromaLocation": loc, "_Matrix": 6}
        clip.release()
        src.release()
