"""Drop-in check at the AviSynth boundary: the B200 plugin and the UNMODIFIED reference plugin (oracle/_ref, opt=0)
are loaded into the same mini-host and driven with the same script calls; frames must agree within the parity bar
and frame properties must propagate.  Where oracle/_ref is absent the CPU oracle stands in for the reference."""
import numpy as np
import pytest

from common import SMALL_CASES, assert_plane_close, make_planes, oracle_frame
from oracle import ref as oref

pytestmark = pytest.mark.gpu

ALIAS = {3: "Jinc36Resize", 4: "Jinc64Resize", 6: "Jinc144Resize", 8: "Jinc256Resize"}


@pytest.fixture(scope="module")
def envs(native_built):
    from jinc_b200 import capi, paths
    from minihost import avs_host as ah

    assert capi.device_count() >= 1, "no CUDA device: the product has no CPU fallback"
    ours = ah.Env()
    ours.load_plugin(paths.b200_plugin())
    theirs = None
    if oref.available():
        theirs = ah.Env()
        theirs.load_plugin(oref.REF_PLUGIN)
    return ours, theirs


@pytest.mark.parametrize("case", SMALL_CASES, ids=[c[0] for c in SMALL_CASES])
def test_plugin_frames_match_reference_plugin(envs, case):
    ours, theirs = envs
    name, fmt, w, h, tw, th, kw = case
    frames = [make_planes(fmt, w, h, "noise", seed=s) for s in range(3)]
    # use the alias function when the case can be expressed through it (no blur), as a script would
    kw2 = dict(kw)
    fn = "JincResize"
    if "blur" not in kw2 and kw2.get("tap", 3) in ALIAS:
        fn = ALIAS[kw2.pop("tap", 3)]
    src = ours.source(fmt, w, h, frames)
    clip = ours.invoke(fn, src, tw, th, **kw2)
    assert clip.size == (tw, th)
    assert clip.mt_mode == 1
    rclip = rsrc = None
    if theirs is not None:
        rsrc = theirs.source(fmt, w, h, frames)
        rclip = theirs.invoke("JincResize", rsrc, tw, th, opt=0, **kw)
    for n in range(3):
        got, props = clip.get_frame(n)
        if rclip is not None:
            ref, _ = rclip.get_frame(n)
        else:
            ref, _ = oracle_frame(fmt, w, h, tw, th, frames[n], **kw)
        for i, (g, r) in enumerate(zip(got, ref)):
            assert_plane_close(g, r, fmt.bits == 32, f"{name}/frame{n}/plane{i}")
        if fmt.family in ("420", "422", "411", "yuva420", "yuva422"):
            want = {"mpeg2": 0, "mpeg1": 1, "topleft": 2}[kw.get("cplace", "mpeg2").lower()]
            assert props["_ChromaLocation"] == want
        else:
            assert "_ChromaLocation" not in props
    for c in (clip, src, rclip, rsrc):
        if c is not None:
            c.release()


def test_cplace_defaults_from_frame_property(envs):
    from minihost import avs_host as ah

    ours, theirs = envs
    fmt, w, h = ah.YV12, 96, 64
    planes = make_planes(fmt, w, h)
    for loc, name in ((0, "mpeg2"), (1, "mpeg1"), (2, "topleft")):
        src = ours.source(fmt, w, h, [planes], props={"_ChromaLocation": loc})
        got, props = ours.invoke("Jinc36Resize", src, 192, 128).get_frame(0)
        ref, _ = oracle_frame(fmt, w, h, 192, 128, planes, tap=3, cplace=name)
        for g, r in zip(got, ref):
            assert_plane_close(g, r, False, f"cplace-from-prop/{name}")
        assert props["_ChromaLocation"] == loc


def test_prefetch_threads_share_one_instance(envs):
    """Frame-parallel get_frame from 6 host threads on ONE instance (MT_NICE_FILTER): every frame still correct."""
    from minihost import avs_host as ah

    ours, _ = envs
    fmt, w, h = ah.YV12, 160, 90
    frames = [make_planes(fmt, w, h, "noise", seed=s) for s in range(4)]
    src = ours.source(fmt, w, h, frames, num_frames=64)
    clip = ours.invoke("Jinc36Resize", src, 320, 180)
    assert clip.pull(0, 48, threads=6) > 0
    for n in (0, 5, 10, 63):
        got, _ = clip.get_frame(n)
        ref, _ = oracle_frame(fmt, w, h, 320, 180, frames[n % 4], tap=3)
        for g, r in zip(got, ref):
            assert_plane_close(g, r, False, f"mt/frame{n}")
    clip.release()
    src.release()


def test_no_frames_or_clips_leak(envs):
    from minihost import avs_host as ah

    ours, _ = envs
    fmt, w, h = ah.YV12, 64, 64
    f0, c0 = ah.Env.live_objects()
    src = ours.source(fmt, w, h, [make_planes(fmt, w, h)])
    clip = ours.invoke("Jinc36Resize", src, 128, 128)  # cplace taken from frame 0: that frame must be released
    clip.get_frame(0)
    clip.release()
    src.release()
    assert ah.Env.live_objects() == (f0, c0)
