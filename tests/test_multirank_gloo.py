"""N > 1 path on CPU: two gloo ranks, each driving its own independent lookahead stream (through the CPU
sim-engine), then the same barrier / max-over-ranks / sum-of-frames plumbing bench.py uses under NCCL."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, simdir, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    for p in (ROOT, os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import importlib.util
    import _pkg
    pkg = _pkg.load_pkg(); synth = _pkg.load_synth()
    spec = importlib.util.spec_from_file_location("shard", os.path.join(ROOT, "x265-amod_b200", "shard.py"))
    shard = importlib.util.module_from_spec(spec); spec.loader.exec_module(shard)
    dist = shard.init("gloo")
    r, w, _ = shard.rank_info()
    mine = shard.streams_for_rank(4, r, w)
    frames = 0
    types = {}
    for s in mine:
        seq = synth.SynthSequence(176, 144, depth=8, seed=10 + s, cuts=(9,))
        la = pkg.Lookahead(176, 144, depth=8, lib_path=os.path.join(simdir, "libx265la_sim8.so"), bframes=3, lookaheadDepth=8)
        out = pkg.run_sequence(la, (seq.frame(i) for i in range(16)), collect=False)
        la.close()
        frames += len(out)
        types[s] = [f["sliceType"] for f in out]
    dist.barrier()
    fps, ms = shard.aggregate_throughput(dist, frames, 100.0 + 50.0 * r)
    q.put((r, mine, frames, fps, ms, types))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_streams(simdir):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, simdir, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, f0, fps0, ms0, t0), (r1, s1, f1, fps1, ms1, t1) = res
    assert s0 == [0, 2] and s1 == [1, 3]                 # every stream owned by exactly one rank
    assert f0 == 32 and f1 == 32
    assert ms0 == ms1 == 150.0                           # max over ranks
    assert abs(fps0 - 64 / 0.150) < 1e-6 and fps0 == fps1  # total frames / slowest rank
    # streams are independent: a single-rank run of stream 1 gives the same decisions
    sys.path.insert(0, ROOT)
    import _pkg
    pkg = _pkg.load_pkg(); synth = _pkg.load_synth()
    seq = synth.SynthSequence(176, 144, depth=8, seed=11, cuts=(9,))
    la = pkg.Lookahead(176, 144, depth=8, lib_path=os.path.join(simdir, "libx265la_sim8.so"), bframes=3, lookaheadDepth=8)
    out = pkg.run_sequence(la, (seq.frame(i) for i in range(16)), collect=False)
    assert [f["sliceType"] for f in out] == t1[1]


def _shard_worker(rank, world, port, simdir, q, case_name, extra):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    for p in (ROOT, os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import importlib.util
    import _pkg
    import cases
    pkg = _pkg.load_pkg(); synth = _pkg.load_synth()
    spec = importlib.util.spec_from_file_location("shard", os.path.join(ROOT, "x265-amod_b200", "shard.py"))
    shard = importlib.util.module_from_spec(spec); spec.loader.exec_module(shard)
    dist = shard.init("gloo")
    case = cases.get_case(case_name)
    name, depth, w, h, n, skw, rkw = case
    seq = cases.make_seq(synth, case)
    kw = cases.la_kwargs(rkw); kw.update(extra)
    la = pkg.Lookahead(w, h, depth=depth, lib_path=os.path.join(simdir, "libx265la_sim%d.so" % depth), shardCount=world, **kw)
    la.shard(rank, world, shard.make_exchange(dist, pkg.EXCHANGE_FN, cuda=False))
    out = pkg.run_sequence(la, (seq.frame(i) for i in range(n)), planes=False)
    cnt = pkg.Counters()
    la.lib.x265cu_get_counters.argtypes = [__import__("ctypes").c_void_p, __import__("ctypes").POINTER(pkg.Counters)]
    la.lib.x265cu_get_counters(la.engine(), __import__("ctypes").byref(cnt))
    la.close()
    dist.barrier()
    import pickle
    q.put((rank, pickle.dumps(out), int(cnt.search_jobs), int(cnt.cost_jobs)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case_name,extra,world", [("base8", {}, 2), ("static_noise", dict(asyncDepth=5), 2), ("fade8", dict(speculate=2), 2),
                                                   ("base8", dict(asyncDepth=6), 4), ("fade8", {}, 3), ("hme_golden", {}, 2)])
def test_two_ranks_shard_one_stream(case_name, extra, world, simdir):
    """SURVEY 8e level 2: one stream, searches / estimates split by source frame over two ranks, stores exchanged after
    every batch (gloo here, NCCL on the GPU box).  Both ranks must publish exactly what a single rank publishes (the
    golden fixture), and each must have computed only its share of the jobs."""
    import pickle
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases, compare, golden_io
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, simdir, q, case_name, extra)) for r in range(world)]
    for p in procs:
        p.start()
    res = []
    import queue as _queue
    import time as _time
    t0 = _time.time()
    while len(res) < len(procs):
        try:
            res.append(q.get(timeout=2))
        except _queue.Empty:
            dead = [p.exitcode for p in procs if not p.is_alive() and p.exitcode]
            if dead or _time.time() - t0 > 600:       # a rank failed: do not wait for the collective's timeout on the others
                for p in procs:
                    if p.is_alive():
                        p.kill()
                raise AssertionError("a rank failed (exit codes %s)" % dead)
    res.sort()
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    case = cases.get_case(case_name)
    want = golden_io.load(case_name)
    jobs = []
    for rank, blob, nsearch, ncost in res:
        got = pickle.loads(blob)
        bad = compare.compare_runs(want, got, check_planes=False, cutree=case[6].get("cuTree", 1), weightp=case[6].get("weightp", 1),
                                   decision_rank=rank == 0)
        assert not bad, "rank %d:\n%s" % (rank, "\n".join(bad[:10]))
        jobs.append((nsearch, ncost))
    # the work really was split: no rank did (nearly) all of it
    tot_s = sum(j[0] for j in jobs); tot_c = sum(j[1] for j in jobs)
    assert min(j[0] for j in jobs) > 0.6 / world * tot_s and min(j[1] for j in jobs) > 0.6 / world * tot_c, jobs
