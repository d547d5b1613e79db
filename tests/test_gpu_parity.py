"""GPU suite (-m gpu): the CUDA lookahead, called through the C ABI / the host Lookahead, against
(1) the C oracle at primitive level, (2) the committed golden fixtures of the unmodified reference,
(3) the live reference build (oracle/_ref) when its libraries travelled with the snapshot, and
(4) the oracle-backed host pipeline for every configuration, incl. full-size property checks."""
import ctypes as C
import os

import numpy as np
import pytest

import cases
import compare
import golden_io
import refbind

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cmp(want, got, case, planes=True):
    rkw = case[6]
    return compare.compare_runs(want, got, check_planes=planes, cutree=rkw.get("cuTree", 1), weightp=rkw.get("weightp", 1),
                                vbv=bool(rkw.get("vbvBufferSize")), skip_propagate=tuple(cases.FORCED.get(case[0], {})) + tuple(cases.PASS2.get(case[0], {})))


def _sim(simdir, depth):
    return os.path.join(simdir, "libx265la_sim%d.so" % depth)


def _as_ref_layout(frames):
    for f in frames:
        f["mvs"][~f["searched"]] = 0
        f["mvs"][~f["searched"], 0, 0] = 0x7FFF
    return frames


@pytest.mark.parametrize("depth", [8, 10, 12])
def test_block_metrics_vs_oracle(depth, pkg, simdir):
    """T1: SAD/SATD kernels on the reference's pixelharness recipe (random / min / max buffers)"""
    la = pkg.Lookahead(320, 192, depth=depth)
    eng = pkg.load_engine()
    eng.x265cu_debug_block_metrics.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    orc = C.CDLL(os.path.join(simdir, "liboracle%d.so" % depth))
    dt = np.uint8 if depth == 8 else np.uint16
    maxv = (1 << depth) - 1
    rng = np.random.default_rng(5)
    n = 4096
    a = rng.integers(0, maxv + 1, (n, 64)).astype(dt)
    b = rng.integers(0, maxv + 1, (n, 64)).astype(dt)
    a[:64] = 0; b[:64] = maxv; a[64:128] = maxv; b[64:128] = 0; b[128:192] = a[128:192]
    b[192:1024] = np.clip(a[192:1024].astype(np.int64) + rng.integers(-3, 4, (832, 64)), 0, maxv).astype(dt)
    sad = np.zeros(n, np.int32); satd = np.zeros(n, np.int32)
    assert eng.x265cu_debug_block_metrics(la.engine(), a.ctypes.data, b.ctypes.data, n, sad.ctypes.data, satd.ctypes.data) == 0
    for i in range(n):
        pa = a[i].ctypes.data_as(C.c_void_p); pb = b[i].ctypes.data_as(C.c_void_p)
        assert sad[i] == orc.or_sad8x8(pa, 8, pb, 8), i
        assert satd[i] == orc.or_satd8x8(pa, 8, pb, 8), i
    la.close()


@pytest.mark.parametrize("depth", [8, 10, 12])
def test_tiled_motion_compensation_vs_oracle(depth, pkg, synth, simdir):
    """T1: the tiled row fetch + lowresMC + SAD/SATD (lowresQPelCost, lowres.h:98-124) for random blocks and
    quarter-pel vectors, including vectors that reach into the plane margins"""
    w, h = 328, 184
    la = pkg.Lookahead(w, h, depth=depth)
    eng = pkg.load_engine()
    seq = synth.SynthSequence(w, h, depth=depth, seed=9)
    la.add_picture(*seq.frame(0)); la.add_picture(*seq.frame(3))
    ctx = la.engine()
    g = la.geom
    dt = np.uint8 if depth == 8 else np.uint16
    planes = []
    eng.x265cu_fetch_frame.argtypes = [C.c_void_p, C.c_int32, C.POINTER(pkg.FrameOut)]
    for slot in (0, 1):
        pl = np.zeros((4, g.plane_lines, g.stride), dt)
        fo = pkg.FrameOut(); fo.planes = pl.ctypes.data
        assert eng.x265cu_fetch_frame(ctx, slot, C.byref(fo)) == 0
        planes.append(pl)
    orc = C.CDLL(os.path.join(simdir, "liboracle%d.so" % depth))
    rng = np.random.default_rng(11)
    n = 2048
    cu = rng.integers(0, g.ncu, n).astype(np.int32)
    mv = rng.integers(-60, 61, (n, 2)).astype(np.int32)
    mv[:64] = 0
    # vectors to the allowed extremes: frame edge +- 8 full-pel, plus the hexagon overshoot
    for i in range(64, 192):
        cx, cy = cu[i] % g.bw, cu[i] // g.bw
        mv[i, 0] = 4 * rng.choice([-cx * 8 - 12, (g.bw - cx - 1) * 8 + 12]) + rng.integers(0, 4)
        mv[i, 1] = 4 * rng.choice([-cy * 8 - 11, (g.bh - cy - 1) * 8 + 11]) + rng.integers(0, 4)
    sad = np.zeros(n, np.int32); satd = np.zeros(n, np.int32)
    eng.x265cu_debug_mc_metrics.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    assert eng.x265cu_debug_mc_metrics(ctx, 1, 0, cu.ctypes.data, mv.ctypes.data, n, sad.ctypes.data, satd.ctypes.data) == 0
    mx, my = g.margin_x, g.margin_y

    def blk(pl, X, Y):
        return pl[Y:Y + 8, X:X + 8].astype(np.int64)
    for i in range(n):
        cx, cy = cu[i] % g.bw, cu[i] // g.bw
        X0, Y0 = mx + 8 * cx, my + 8 * cy
        qx, qy = int(mv[i, 0]), int(mv[i, 1])
        fenc = np.ascontiguousarray(blk(planes[1][0], X0, Y0).astype(dt))
        hA = (qy & 2) | ((qx & 2) >> 1)
        A = blk(planes[0][hA], X0 + (qx >> 2), Y0 + (qy >> 2))
        if (qx | qy) & 1:
            qx2, qy2 = qx + (qx & 1), qy + (qy & 1)
            hB = (qy2 & 2) | ((qx2 & 2) >> 1)
            B = blk(planes[0][hB], X0 + (qx2 >> 2), Y0 + (qy2 >> 2))
            A = (A + B + 1) >> 1
        pred = np.ascontiguousarray(A.astype(dt))
        assert sad[i] == int(np.abs(fenc.astype(np.int64) - A).sum()), (i, cu[i], mv[i])
        assert satd[i] == orc.or_satd8x8(fenc.ctypes.data_as(C.c_void_p), 8, pred.ctypes.data_as(C.c_void_p), 8), (i, cu[i], mv[i])
    la.close()


@pytest.mark.parametrize("name", cases.GOLDEN)
def test_cuda_pipeline_matches_golden(name, pkg, synth):
    case = cases.get_case(name)
    want = golden_io.load(name)
    got = cases.run_ours(pkg, synth, case, planes=True)
    bad = _cmp(want, got, case, planes=False)
    assert not bad, "\n".join(bad[:10])
    for w, g in zip(want, got):
        assert w["planesum"] == golden_io.planesum(g["planes"]), "lowres planes differ at poc %d" % w["poc"]


@pytest.mark.parametrize("name", [c[0] for c in cases.CASES if c[0] not in cases.CPU_ONLY])
def test_cuda_pipeline_matches_reference_or_oracle(name, pkg, synth, simdir):
    case = cases.get_case(name)
    if refbind.available(case[1]):
        want = cases.run_reference(refbind, synth, case)
    else:
        want = _as_ref_layout(cases.run_ours(pkg, synth, case, lib_path=_sim(simdir, case[1])))
    got = cases.run_ours(pkg, synth, case)
    bad = _cmp(want, got, case)
    assert not bad, "\n".join(bad[:10])


def test_cuda_speculation_off_is_identical(pkg, synth):
    case = cases.get_case("base8")
    b = cases.run_ours(pkg, synth, case, speculate=0, planes=False)
    bad = compare.compare_runs(golden_io.load("base8"), b, check_planes=False)
    assert not bad, "\n".join(bad[:10])


@pytest.mark.parametrize("name", ["base8", "static_noise", "fade10", "pool16", "b8"])
@pytest.mark.parametrize("mode", [(2, 9), (2, 0), (1, 12), (1, 3), (0, 0)])
def test_cuda_speculation_modes_and_async_depth(name, mode, pkg, synth, simdir):
    """streaming (default) / per-decision / on-demand scheduling of the GPU batches and extra input delay:
    same published Lowres state, bit for bit"""
    case = cases.get_case(name)
    if refbind.available(case[1]):
        want = cases.run_reference(refbind, synth, case, planes=False)
    else:
        want = _as_ref_layout(cases.run_ours(pkg, synth, case, lib_path=_sim(simdir, case[1]), planes=False))
    got = cases.run_ours(pkg, synth, case, planes=False, speculate=mode[0], asyncDepth=mode[1])
    bad = compare.compare_runs(want, got, check_planes=False, cutree=case[6].get("cuTree", 1), weightp=case[6].get("weightp", 1))
    assert not bad, "\n".join(bad[:10])


@pytest.mark.parametrize("name", [c[0] for c in cases.FULL_SIZE])
def test_baseline_configs_at_full_size_match_live_reference(name, pkg, synth):
    """BASELINE.json's configurations themselves (2160p main10 rc-lookahead 60 bframes 8 -- the benchmarked one --, 1080p
    8-bit rc-lookahead 20 bframes 4, 1080p preset-slower depths with weightp on fades / cuts / flashes), full size, against
    the live unmodified reference: every published field of every decided frame, bit-exact."""
    case = [c for c in cases.FULL_SIZE if c[0] == name][0]
    if not refbind.available(case[1]):
        pytest.skip("oracle/_ref not in this snapshot")
    bad, n = cases.compare_streaming(pkg, synth, refbind, compare, case, asyncDepth=16)
    assert not bad, "\n".join(bad[:10])
    assert n == case[4]


def test_config4_4320p_matches_c_oracle(pkg, synth, simdir):
    """7680x4320 8-bit rc-lookahead 80 bframes 4 (BASELINE configs[3]) on a dozen frames against the C oracle"""
    name, depth, w, h, n, skw, rkw = cases.FULL_SIZE_ORACLE
    seq = synth.SynthSequence(w, h, depth=depth, **skw)
    frames = [seq.frame(i) for i in range(n)]
    kw = cases.la_kwargs(rkw)
    want = _as_ref_layout(pkg.run_sequence(pkg.Lookahead(w, h, depth=depth, lib_path=_sim(simdir, depth), **kw), iter(frames), planes=False))
    got = pkg.run_sequence(pkg.Lookahead(w, h, depth=depth, asyncDepth=8, **kw), iter(frames), planes=False)
    bad = compare.compare_runs(want, got, check_planes=False)
    assert not bad, "\n".join(bad[:10])


@pytest.mark.parametrize("depth,w,h", [(8, 1920, 1080), (10, 3840, 2160), (8, 7680, 4320)])
def test_full_size_properties(depth, w, h, pkg, synth, simdir):
    """BASELINE sizes: oracle spot-check of whole frames plus size-independent properties
    (determinism across runs and across speculation on/off; intra cost independent of neighbours)."""
    n = 14
    seq = synth.SynthSequence(w, h, depth=depth, seed=2, cuts=(7,), n_rects=4)
    frames = [seq.frame(i) for i in range(n)]
    kw = dict(bframes=3, lookaheadDepth=8)

    def run(**extra):
        la = pkg.Lookahead(w, h, depth=depth, **dict(kw, **extra))
        out = pkg.run_sequence(la, iter(frames), planes=False)
        la.close()
        return out
    a = run()
    b = run(speculate=0)
    c = run(asyncDepth=6)
    assert [f["sliceType"] for f in a] == [f["sliceType"] for f in b] == [f["sliceType"] for f in c]
    for x, y in list(zip(a, b)) + list(zip(a, c)):
        assert np.array_equal(x["costEst"], y["costEst"]) and np.array_equal(x["costEstAq"], y["costEstAq"])
        assert np.array_equal(x["intraCost"], y["intraCost"]) and np.array_equal(x["mvs"], y["mvs"])
        assert np.array_equal(x["qpCuTreeOffset"], y["qpCuTreeOffset"])
    # full-size oracle check through the sim engine (the C oracle finishes 14 frames in seconds at 1080p;
    # at 2160p only when the run is affordable)
    if w <= 1920:
        want = _as_ref_layout(pkg.run_sequence(pkg.Lookahead(w, h, depth=depth, lib_path=_sim(simdir, depth), **kw), iter(frames), planes=False))
        bad = compare.compare_runs(want, a, check_planes=False)
        assert not bad, "\n".join(bad[:10])


def _gpu_shard_worker(rank, world, port, q, case_name):
    import sys
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests")):
        sys.path.insert(0, p)
    import importlib.util, pickle
    import torch
    import _pkg
    import cases as cs
    torch.cuda.set_device(rank)
    pkg_ = _pkg.load_pkg(); synth_ = _pkg.load_synth()
    spec = importlib.util.spec_from_file_location("shard", os.path.join(root, "x265-amod_b200", "shard.py"))
    shard = importlib.util.module_from_spec(spec); spec.loader.exec_module(shard)
    dist = shard.init("nccl")
    case = cs.get_case(case_name)
    name, depth, w, h, n, skw, rkw = case
    seq = cs.make_seq(synth_, case)
    la = pkg_.Lookahead(w, h, depth=depth, shardCount=world, device=rank, asyncDepth=6, **cs.la_kwargs(rkw))
    la.shard(rank, world, shard.make_exchange(dist, pkg_.EXCHANGE_FN, cuda=True))
    out = pkg_.run_sequence(la, (seq.frame(i) for i in range(n)), planes=False)
    la.close()
    dist.barrier()
    q.put((rank, pickle.dumps(out)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case_name", ["base8", "static_noise", "fade8"])
def test_cuda_one_stream_sharded_over_two_gpus(case_name, pkg, synth):
    """SURVEY 8e level 2 on hardware: two ranks, two GPUs, NCCL broadcasts of the stores after every batch; both ranks
    must publish the golden / reference state bit for bit."""
    import pickle
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gpu_shard_worker, args=(r, 2, port, q, case_name)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    case = cases.get_case(case_name)
    if refbind.available(case[1]):
        # the shard workers run no getEstimatedPictureCost (rate control belongs to the decision rank's caller): the same here
        want = cases.run_reference(refbind, synth, case, planes=False, estimate=False)
    else:
        want = golden_io.load(case_name)
    for rank, blob in res:
        got = pickle.loads(blob)
        bad = compare.compare_runs(want, got, check_planes=False, cutree=case[6].get("cuTree", 1), weightp=case[6].get("weightp", 1),
                                   decision_rank=rank == 0)
        assert not bad, "rank %d:\n%s" % (rank, "\n".join(bad[:10]))


@pytest.mark.parametrize("nframes,extra", [(0, {}), (1, {}), (2, dict(asyncDepth=4)), (5, dict(asyncDepth=4)), (5, dict(speculate=2))])
def test_cuda_tiny_sequences(nframes, extra, pkg, synth, simdir):
    """empty input and sequences shorter than a mini-GOP / the lookahead / the async depth"""
    case = ("tiny", 8, 176, 144, nframes, dict(cuts=()), dict(bframes=3, lookaheadDepth=10))
    got = cases.run_ours(pkg, synth, case, planes=False, **extra)
    assert len(got) == nframes
    assert sorted(g["poc"] for g in got) == list(range(nframes))
    if nframes:
        if refbind.available(8):
            want = cases.run_reference(refbind, synth, case, planes=False)
        else:
            want = _as_ref_layout(cases.run_ours(pkg, synth, case, lib_path=_sim(simdir, 8), planes=False))
        bad = compare.compare_runs(want, got, check_planes=False)
        assert not bad, "\n".join(bad[:10])


_ASSUMED_WEIGHTS_SCRIPT = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests")); sys.path.insert(0, os.path.join({root!r}, "oracle"))
import _pkg, cases, compare, golden_io, refbind
pkg = _pkg.load_pkg(); synth = _pkg.load_synth()
case = cases.get_case({name!r})
want = cases.run_reference(refbind, synth, case, planes=False) if refbind.available(case[1]) else golden_io.load({name!r})
got = cases.run_ours(pkg, synth, case, planes=False, speculate={spec}, asyncDepth={depth})
bad = compare.compare_runs(want, got, check_planes=False, cutree=case[6].get("cuTree", 1), weightp=case[6].get("weightp", 1))
print("\n".join(bad[:10]))
sys.exit(1 if bad else 0)
"""


@pytest.mark.parametrize("name", ["fade8", "fade10", "fade_weightb_pool"])
@pytest.mark.parametrize("mode", [(1, 12), (2, 6)])
def test_cuda_weights_assumed_then_redone(name, mode):
    """Batches are cut without waiting for the pixel sums weightp needs: a pair whose sums are still in flight is enqueued
    unweighted and settled later (Lookahead::verifyWeights), its searches and costs redone on the weighted reference when the
    analysis asks for weights.  X265CU_FRAME_READY_NEVER makes every frame look in flight, so the fade sequences take that path
    for every pair (a fresh process: the engine reads the variable once)."""
    import subprocess
    import sys
    if name != "fade8" and not refbind.available(cases.get_case(name)[1]):
        pytest.skip("needs the live reference")
    env = dict(os.environ, X265CU_FRAME_READY_NEVER="1", X265LA_DEBUG_WEIGHTS="1")
    r = subprocess.run([sys.executable, "-c", _ASSUMED_WEIGHTS_SCRIPT.format(root=ROOT, name=name, spec=mode[0], depth=mode[1])],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr[-2000:]
    assert "verifyWeights: redo" in r.stderr, "the fade never took the redo path"
