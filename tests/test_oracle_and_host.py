"""CPU suite (no GPU): the C oracle + the PRODUCT's host decision logic, run through the CPU
sim-engine (tests/simengine), must reproduce the unmodified reference bit for bit -- against the
committed golden fixtures always, and against the live reference build (oracle/_ref) when present."""
import os

import numpy as np
import pytest

import cases
import compare
import golden_io
import refbind


def _sim(simdir, depth):
    import os
    return os.path.join(simdir, "libx265la_sim%d.so" % depth)


def _cmp(want, got, case):
    rkw = case[6]
    return compare.compare_runs(want, got, check_planes=True, cutree=rkw.get("cuTree", 1), weightp=rkw.get("weightp", 1),
                                vbv=bool(rkw.get("vbvBufferSize")), skip_propagate=tuple(cases.FORCED.get(case[0], {})) + tuple(cases.PASS2.get(case[0], {})))


@pytest.mark.parametrize("name", cases.GOLDEN)
def test_sim_pipeline_matches_golden(name, pkg, synth, simdir):
    case = cases.get_case(name)
    want = golden_io.load(name)
    got = cases.run_ours(pkg, synth, case, lib_path=_sim(simdir, case[1]), planes=True)
    bad = _cmp(want, got, case)
    assert not bad, "\n".join(bad[:10])
    for w, g in zip(want, got):
        assert w["planesum"] == golden_io.planesum(g["planes"]), "lowres planes differ at poc %d" % w["poc"]


@pytest.mark.parametrize("name", [c[0] for c in cases.CASES])
def test_sim_pipeline_matches_live_reference(name, pkg, synth, simdir):
    case = cases.get_case(name)
    if not refbind.available(case[1]):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    want = cases.run_reference(refbind, synth, case)
    got = cases.run_ours(pkg, synth, case, lib_path=_sim(simdir, case[1]))
    bad = _cmp(want, got, case)
    assert not bad, "\n".join(bad[:10])


@pytest.mark.parametrize("name", ["hme_default", "hme_hexhex10", "hme_umhdia_pool_fade", "hme_star", "hme_fullhex"])
def test_generic_search_source_on_cpu(name):
    """--hme: the source text of the product's dia / hex / umh searches (csrc/la_me_generic.cuh, what search_hme_kernel compiles)
    built for the CPU with a scalar evaluator reproduces the live reference (the warp-level evaluators are the GPU suite's job)"""
    import subprocess
    import sys
    if not refbind.available(cases.get_case(name)[1]):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "hme_generic_check.py"), name], env=dict(os.environ, X265SIM_GENERIC_ME="1"),
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]


def test_hme_refusals(pkg, simdir):
    """--hme outside what is built fails loudly at create: star / sea / full at levels 0-1, and active lookahead slices (where the
    reference's two levels race on uninitialised vectors)"""
    sim = _sim(simdir, 8)
    with pytest.raises(RuntimeError, match="hme-search"):
        pkg.Lookahead(960, 544, depth=8, lib_path=sim, bEnableHME=1, hmeSearchMethod=(4, 2))
    with pytest.raises(RuntimeError, match="lookahead slices"):
        pkg.Lookahead(1280, 720, depth=8, lib_path=sim, bEnableHME=1, poolWorkers=8, lookaheadSlices=4)
    with pytest.raises(RuntimeError, match="fades"):
        pkg.Lookahead(320, 192, depth=8, lib_path=sim, aqMode=4, bEnableFades=1)
    # like Encoder::configure (encoder.cpp:3914-3943) the C surface fixes bframes 7 / b-adapt 0 for four temporal layers
    la = pkg.Lookahead(320, 192, depth=8, lib_path=sim, bEnableTemporalSubLayers=4, bframes=4)
    assert la.geom.nb == 9
    la.close()
    # below 540 lines the encoder itself turns --hme off (encoder.cpp:4400-4407): accepted, and plain searches run
    pkg.Lookahead(320, 192, depth=8, lib_path=sim, bEnableHME=1).close()


def test_speculation_off_gives_same_results(pkg, synth, simdir):
    case = cases.get_case("base8")
    a = cases.run_ours(pkg, synth, case, lib_path=_sim(simdir, 8), speculate=1)
    b = cases.run_ours(pkg, synth, case, lib_path=_sim(simdir, 8), speculate=0)
    bad = compare.compare_runs(golden_io.load("base8"), b, check_planes=False)
    assert not bad, "\n".join(bad[:10])
    for x, y in zip(a, b):
        assert np.array_equal(x["costEst"], y["costEst"]) and x["sliceType"] == y["sliceType"]


@pytest.mark.parametrize("name", ["base8", "static_noise", "fade8", "pool16", "nob"])
@pytest.mark.parametrize("mode", [(2, 0), (2, 7), (1, 0), (1, 5), (0, 3)])
def test_speculation_modes_and_async_depth_do_not_change_results(name, mode, pkg, synth, simdir):
    """streaming / per-decision / on-demand job scheduling and any extra input delay publish the same Lowres state"""
    case = cases.get_case(name)
    got = cases.run_ours(pkg, synth, case, lib_path=_sim(simdir, case[1]), planes=False, speculate=mode[0], asyncDepth=mode[1])
    bad = compare.compare_runs(golden_io.load(name), got, check_planes=False, cutree=case[6].get("cuTree", 1),
                               weightp=case[6].get("weightp", 1))
    assert not bad, "\n".join(bad[:10])


def test_coverage_of_special_paths(pkg, synth, simdir):
    """the fixtures really exercise weightp and the B-frame zero-MV skip rule"""
    got = cases.run_ours(pkg, synth, cases.get_case("fade8"), lib_path=_sim(simdir, 8), planes=False)
    assert sum(int(np.sum(g["weightState"] == 2)) for g in got) > 0
    got = cases.run_ours(pkg, synth, cases.get_case("static_noise"), lib_path=_sim(simdir, 8), planes=False)
    cheap = sum(int(np.sum((g["mvs"][0, 1:, :, 0] == 0) & (g["mvs"][0, 1:, :, 1] == 0) & (g["mvCosts"][0, 1:] < 64) &
                           (g["mvCosts"][0, 1:] > 0))) for g in got)
    assert cheap > 0
    types = [g["sliceType"] for g in cases.run_ours(pkg, synth, cases.get_case("base8"), lib_path=_sim(simdir, 8), planes=False)]
    assert 2 in types, "scene cut not detected"


@pytest.mark.parametrize("depth", [8, 10, 12])
def test_mvcost_table_and_lambda(depth, pkg, simdir):
    """BitCost's row for the lookahead QP: the oracle's AND the product's table (x265la_mvcost_table) against the reference's own, over
    the whole index range a 4320p search can reach (at 12 bits, lambda 256, a float-logf pipeline is off by one in 15 entries)"""
    if not refbind.available(depth):
        pytest.skip("oracle/_ref not built")
    import ctypes as C, os
    lib = C.CDLL(os.path.join(os.path.dirname(__file__), "_build", "liboracle%d.so" % depth))
    n = 60000
    qp, ref = refbind.mvcost_table(depth, n)
    tab = np.zeros(2 * n + 1, np.uint16)
    lib.or_build_mvcost(tab.ctypes.data_as(C.c_void_p), n)
    assert np.array_equal(tab, ref)
    host = C.CDLL(_sim(simdir, depth))      # the product's host library (linked against the sim engine here)
    tab2 = np.zeros(2 * n + 1, np.uint16)
    host.x265la_mvcost_table(depth, n, tab2.ctypes.data_as(C.c_void_p))
    assert np.array_equal(tab2, ref)
    assert lib.or_lookahead_lambda() == refbind.load(depth).ref_lookahead_lambda()
    assert qp == 12 + 6 * (depth - 8)


@pytest.mark.parametrize("depth", [8, 10, 12])
def test_oracle_block_primitives_vs_reference(depth, simdir):
    """the reference's own pixelharness recipe: random, all-min and all-max buffers"""
    if not refbind.available(depth):
        pytest.skip("oracle/_ref not built")
    import ctypes as C, os
    lib = C.CDLL(os.path.join(simdir, "liboracle%d.so" % depth))
    ref = refbind.load(depth)
    dt = np.uint8 if depth == 8 else np.uint16
    maxv = (1 << depth) - 1
    rng = np.random.default_rng(3)
    bufs = [rng.integers(0, maxv + 1, (64, 64)).astype(dt), np.zeros((64, 64), dt), np.full((64, 64), maxv, dt)]
    for a in bufs:
        for b in bufs:
            for (ox, oy) in ((0, 0), (3, 5), (17, 1)):
                pa = a[oy:, ox:]; pb = b[1:, 2:]
                args = (pa.ctypes.data_as(C.c_void_p), 64, pb.ctypes.data_as(C.c_void_p), 64)
                assert lib.or_sad8x8(*args) == ref.ref_sad8x8(*args)
                assert lib.or_satd8x8(*args) == ref.ref_satd8x8(*args)
    lib.or_exp2fix8.argtypes = [C.c_double]
    for x in np.linspace(-60, 60, 4001):
        assert lib.or_exp2fix8(float(x)) == ref.ref_exp2fix8(float(x))


def test_libraries_export_the_declared_abi():
    """libx265cu.so loads without a GPU and exports every symbol include/x265cu.h declares"""
    import ctypes as C, os, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "x265cu.h")).read()
    names = sorted(set(re.findall(r"\b(x265cu_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 25
    libp = os.path.join(root, "x265-amod_b200", "lib", "libx265cu.so")
    if not os.path.exists(libp):
        import __graft_entry__ as ge
        ge.build_product()
    lib = C.CDLL(libp)
    for n in names:
        assert hasattr(lib, n), n
    la = C.CDLL(os.path.join(root, "x265-amod_b200", "lib", "libx265la.so"))
    cap = open(os.path.join(root, "x265-amod_b200", "host", "la_capi.h")).read()
    for n in sorted(set(re.findall(r"\b(x265la_[a-z0-9_]+)\s*\(", cap))):
        assert hasattr(la, n), n


def test_no_cpu_fallback_without_gpu(pkg):
    """without a CUDA device the product fails loudly instead of computing on the CPU"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError) as e:
        pkg.Lookahead(320, 192, depth=8)
    assert "no usable CUDA device" in str(e.value) or "x265cu_create" in str(e.value)


@pytest.mark.parametrize("nframes", [0, 1, 2, 5])
def test_tiny_sequences(nframes, pkg, synth, simdir):
    """empty input and sequences shorter than a mini-GOP / the lookahead: nothing hangs, every frame comes out once, and
    what comes out matches the reference"""
    case = ("tiny", 8, 176, 144, nframes, dict(cuts=()), dict(bframes=3, lookaheadDepth=10))
    got = cases.run_ours(pkg, synth, case, lib_path=_sim(simdir, 8), planes=False)
    assert len(got) == nframes
    assert sorted(g["poc"] for g in got) == list(range(nframes))
    if nframes and refbind.available(8):
        want = cases.run_reference(refbind, synth, case, planes=False)
        bad = compare.compare_runs(want, got, check_planes=False)
        assert not bad, "\n".join(bad[:10])


def test_lookahead_slices_rules(pkg, synth, simdir):
    """--lookahead-slices is ignored without a pool or below 720 lines, like the reference (slicetype.cpp:1035-1045)"""
    case = cases.get_case("base8")
    for extra in (dict(lookaheadSlices=4), dict(lookaheadSlices=4, poolWorkers=4, bFrameAdaptive=0)):
        got = cases.run_ours(pkg, synth, case, lib_path=_sim(simdir, 8), planes=False, **extra)
        if "poolWorkers" not in extra:
            bad = compare.compare_runs(golden_io.load("base8"), got, check_planes=False)
            assert not bad, "\n".join(bad[:10])


@pytest.mark.parametrize("depth", [8, 10])
def test_intrinsics_baseline_shims_are_bit_exact(depth):
    """oracle/ref_simd.cpp (the asm-class CPU baseline): every SSE4.1 shim against the reference's C primitive on the
    pixelharness buffer recipe, and a whole lookahead run with the shims installed against one without"""
    if not refbind.available(depth):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    assert refbind.simd_selftest(depth, 6000) == 0
    import _pkg
    synth = _pkg.load_synth()
    case = cases.get_case("fade8" if depth == 8 else "pool16_10bit")
    try:
        plain = cases.run_reference(refbind, synth, case)
        assert refbind.simd_install(depth, True)
        shim = cases.run_reference(refbind, synth, case)
    finally:
        refbind.simd_install(depth, False)
    for f in shim:      # give the second reference run the key our own runs carry
        f["searched"] = f["mvs"][:, :, 0, 0] != 0x7FFF
    bad = compare.compare_runs(plain, shim, check_planes=True)
    assert not bad, "\n".join(bad[:10])


@pytest.mark.parametrize("name", ["fade8", "fade10", "fade_weightb_pool"])
@pytest.mark.parametrize("mode", [(1, 12), (2, 6), (1, 0)])
def test_weights_assumed_then_redone(name, mode, pkg, synth, simdir, monkeypatch):
    """Lookahead::verifyWeights: with every frame's pixel sums reported "still in flight" all weightp pairs are enqueued
    unweighted, settled when a decision needs them, and the fades' searches and costs redone on the weighted reference"""
    monkeypatch.setenv("X265CU_FRAME_READY_NEVER", "1")
    case = cases.get_case(name)
    if not refbind.available(case[1]):
        pytest.skip("needs the live reference")
    want = cases.run_reference(refbind, synth, case, planes=False)
    got = cases.run_ours(pkg, synth, case, lib_path=os.path.join(simdir, "libx265la_sim%d.so" % case[1]), planes=False,
                         speculate=mode[0], asyncDepth=mode[1])
    bad = compare.compare_runs(want, got, check_planes=False, cutree=case[6].get("cuTree", 1), weightp=case[6].get("weightp", 1))
    assert not bad, "\n".join(bad[:10])
