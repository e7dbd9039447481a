"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/jinc_b200.h declares, the
product fails loudly without a GPU, the host-side LUT agrees with the oracle, and the AviSynth plugin registers the
reference's five functions with identical parameter strings, argument forwarding and error texts."""
import os
import re

import numpy as np
import pytest

from oracle import ref as oref

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi(native_built):
    from jinc_b200 import capi as c
    from jinc_b200 import paths

    if not os.path.exists(paths.cuda_lib()):
        pytest.fail("libjinc_b200.so missing: run __graft_entry__.build()")
    return c


def test_library_exports_every_declared_symbol(capi):
    hdr = open(os.path.join(REPO, "include", "jinc_b200.h")).read()
    declared = set(re.findall(r"JINC_API\s+[\w\s\*]+?\b(jinc_\w+)\s*\(", hdr))
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    lib = capi.lib()
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.jinc_abi_version() == 2


def test_no_oracle_on_the_product_path():
    """Nothing under the package may import, link or execute oracle/."""
    pkg = os.path.join(REPO, "avisynth-jincresize_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", ".cuh")):
                text = open(os.path.join(root, f), errors="ignore").read()
                assert "jinc_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_host_lut_equals_oracle_lut_as_float(capi):
    from oracle import cpu as oc

    for tap in range(1, 17):
        assert capi.radius_for_tap(tap) == oc.radius_for_tap(tap)
        for blur in (0.0, 0.85, 1.0, 1.3):
            a = capi.lut_build(tap, blur).astype(np.float32)
            b = oc.make_lut(tap, blur).astype(np.float32)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (tap, blur)
    assert capi.radius_for_tap(0) == 0.0 and capi.radius_for_tap(17) == 0.0


@pytest.mark.skipif(os.environ.get("JINC_EXPECT_GPU") == "1", reason="GPU box")
def test_compute_entry_points_fail_loudly_without_a_gpu(capi):
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.JincError, match="no CUDA device"):
        capi.Context(0)
    with pytest.raises(capi.JincError, match="no CUDA device"):
        capi.Filter(src_w=64, src_h=64, target_w=128, target_h=128, n_planes=1, sample_bytes=1, bits=8)


PARAMS_FULL = ("cii[src_left]f[src_top]f[src_width]f[src_height]f[quant_x]i[quant_y]i[tap]i[blur]f[cplace]s[threads]i[opt]i"
               "[initial_capacity]i[initial_factor]f")
PARAMS_ALIAS = "cii[src_left]f[src_top]f[src_width]f[src_height]f[quant_x]i[quant_y]i[cplace]s[threads]i"


@pytest.fixture(scope="module")
def plugin_env(native_built):
    from minihost import avs_host as ah
    from jinc_b200 import paths

    env = ah.Env()
    assert env.load_plugin(paths.b200_plugin()) == "JincResize"
    return env


def test_plugin_registers_the_reference_functions(plugin_env, have_ref):
    from minihost import avs_host as ah
    from jinc_b200 import paths

    assert plugin_env.function_params("JincResize") == PARAMS_FULL
    for fn in ("Jinc36Resize", "Jinc64Resize", "Jinc144Resize", "Jinc256Resize"):
        assert plugin_env.function_params(fn) == PARAMS_ALIAS
    if have_ref:  # and they are byte-identical to what the reference registers
        renv = ah.Env()
        renv.load_plugin(oref.REF_PLUGIN)
        for fn in ("JincResize", "Jinc36Resize", "Jinc64Resize", "Jinc144Resize", "Jinc256Resize"):
            assert plugin_env.function_params(fn) == renv.function_params(fn)


ERROR_CASES = [
    (dict(tap=0), "JincResize: tap must be between 1..16."),
    (dict(tap=17), "JincResize: tap must be between 1..16."),
    (dict(quant_x=0), "JincResize: quant_x must be between 1..256."),
    (dict(quant_y=257), "JincResize: quant_y must be between 1..256."),
    (dict(cplace="center"), "JincResize: cplace must be MPEG2, MPEG1 or topleft."),
    (dict(opt=4), "JincResize: opt higher than 3 is not allowed."),
    (dict(threads=2), "JincResize: threads must be either 0 or 1."),
    (dict(initial_factor=0.5), "JincResize: initial_factor must be eqaul to or greater than 1.0."),
    (dict(initial_capacity=0), "JincResize: initial_capacity must be greater than 0."),
]


OPT_CASES = [
    # (opt, CPU flags the host reports, expected error or None) -- src/JincResize.cpp:747-756
    (3, 0x2000 | 0x400, "JincResize: opt=3 requires AVX-512F."),
    (2, 0x400, "JincResize: opt=2 requires AVX2."),
    (1, 0x20, "JincResize: opt=1 requires SSE4.1."),
]


@pytest.mark.parametrize("opt,flags,msg", OPT_CASES)
def test_plugin_opt_cpu_feature_errors_match_reference(native_built, have_ref, opt, flags, msg):
    """opt selects nothing on the GPU path, but a script that asks for a SIMD level the host CPU lacks gets the
    reference's error, word for word (checked before any GPU work, so this runs without a GPU)."""
    from minihost import avs_host as ah
    from jinc_b200 import paths

    libs = [paths.b200_plugin()] + ([oref.REF_PLUGIN] if have_ref else [])
    for lib in libs:
        env = ah.Env()
        env.load_plugin(lib)
        env.set_cpu_flags(flags)
        src = env.source(ah.Format("y", 8), 64, 64, [[np.zeros((64, 64), np.uint8)]])
        with pytest.raises(ah.AvsError) as e:
            env.invoke("JincResize", src, 96, 96, opt=opt)
        assert str(e.value) == msg, lib


@pytest.mark.parametrize("kw,msg", ERROR_CASES)
def test_plugin_error_texts_match_reference(plugin_env, have_ref, kw, msg):
    from minihost import avs_host as ah
    from jinc_b200 import paths

    z = np.zeros((64, 64), np.uint8)
    src = plugin_env.source(ah.Format("y", 8), 64, 64, [[z]])
    with pytest.raises(ah.AvsError) as e:
        plugin_env.invoke("JincResize", src, 96, 96, **kw)
    assert str(e.value) == msg
    if have_ref:
        renv = ah.Env()
        renv.load_plugin(oref.REF_PLUGIN)
        rsrc = renv.source(ah.Format("y", 8), 64, 64, [[z]])
        with pytest.raises(ah.AvsError) as r:
            renv.invoke("JincResize", rsrc, 96, 96, **kw)
        assert str(r.value) == msg


def test_plugin_rejects_topleft_outside_420_and_packed_and_old_hosts(plugin_env):
    from minihost import avs_host as ah
    from jinc_b200 import paths

    f422 = ah.Format("422", 8)
    src = plugin_env.source(f422, 64, 64, [[np.zeros((64, 64), np.uint8), np.zeros((64, 32), np.uint8), np.zeros((64, 32), np.uint8)]])
    with pytest.raises(ah.AvsError, match="topleft must be used only for 4:2:0"):
        plugin_env.invoke("JincResize", src, 96, 96, cplace="topleft")
    src2 = plugin_env.source(f422, 64, 64, [[np.zeros((64, 64), np.uint8), np.zeros((64, 32), np.uint8), np.zeros((64, 32), np.uint8)]],
                             props={"_ChromaLocation": 5})
    with pytest.raises(ah.AvsError, match="invalid _ChromaLocation"):
        plugin_env.invoke("JincResize", src2, 96, 96)
    old = ah.Env()
    old.load_plugin(paths.b200_plugin())
    for version, bugfix, ok in ((8, 0, False), (9, 1, False), (9, 2, True), (10, 0, True)):
        old.set_interface(version, bugfix)
        s = old.source(ah.Format("y", 8), 64, 64, [[np.zeros((64, 64), np.uint8)]])
        with pytest.raises(ah.AvsError) as e:
            old.invoke("JincResize", s, 96, 96, tap=99)  # a later check fails => the version gate passed
        assert ("r3688" not in str(e.value)) == ok, (version, bugfix, str(e.value))


def test_alias_functions_cannot_pass_blur_or_opt(plugin_env):
    from minihost import avs_host as ah

    src = plugin_env.source(ah.Format("y", 8), 64, 64, [[np.zeros((64, 64), np.uint8)]])
    for fn in ("Jinc36Resize", "Jinc64Resize", "Jinc144Resize", "Jinc256Resize"):
        with pytest.raises(ah.AvsError, match='does not have a named argument "blur"'):
            plugin_env.invoke(fn, src, 96, 96, blur=0.9)
        # and they forward what they do accept: a bad quant_x reaches JincResize's own validation
        with pytest.raises(ah.AvsError, match="quant_x must be between"):
            plugin_env.invoke(fn, src, 96, 96, quant_x=999)
