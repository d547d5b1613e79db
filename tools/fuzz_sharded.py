"""Randomised check of the sharded-stream mode (one stream over N ranks, SURVEY 8e level 2) on CPU: for seeded random configurations
(tools/fuzz_host_vs_reference.random_case) N gloo ranks run the product's host logic over the sim engine with the searches / estimates
split by source frame, and every rank must publish exactly what a single unsharded run publishes.
python tools/fuzz_sharded.py [n_cases] [first_seed] [world]"""
import os
import pickle
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)


def worker(rank, world, port, seed, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import importlib.util
    import _pkg
    import build_sim
    import cases
    import fuzz_host_vs_reference as fz
    pkg = _pkg.load_pkg(); synth = _pkg.load_synth()
    spec = importlib.util.spec_from_file_location("shard", os.path.join(ROOT, "x265-amod_b200", "shard.py"))
    shard = importlib.util.module_from_spec(spec); spec.loader.exec_module(shard)
    dist = shard.init("gloo")
    case = fz.random_case(seed)
    name, depth, w, h, n, skw, rkw = case
    seq = cases.make_seq(synth, case)
    kw = cases.la_kwargs(rkw)
    simdir = os.path.join(ROOT, "tests", "_build")
    try:
        la = pkg.Lookahead(w, h, depth=depth, lib_path=os.path.join(simdir, "libx265la_sim%d.so" % depth), shardCount=world, **kw)
        la.shard(rank, world, shard.make_exchange(dist, pkg.EXCHANGE_FN, cuda=False))
        mono = bool(rkw.get("csp400"))
        out = pkg.run_sequence(la, ((seq.frame(i)[0], None, None) if mono else seq.frame(i) for i in range(n)), planes=False)
        la.close()
        q.put((rank, pickle.dumps(out)))
    except RuntimeError as e:
        q.put((rank, pickle.dumps(str(e))))
    dist.barrier()
    dist.destroy_process_group()


def main():
    import torch.multiprocessing as mp
    import _pkg
    import build_sim
    import cases
    import compare
    import fuzz_host_vs_reference as fz
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    world = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    pkg = _pkg.load_pkg(); synth = _pkg.load_synth()
    simdir = build_sim.build()
    ctx = mp.get_context("spawn")
    bad_cases = refused = 0
    t0 = time.time()
    for seed in range(first, first + n_cases):
        case = fz.random_case(seed)
        name, depth, w, h, n, skw, la = case
        try:
            want = cases.run_ours(pkg, synth, case, lib_path=os.path.join(simdir, "libx265la_sim%d.so" % depth), planes=False)
        except RuntimeError:
            refused += 1
            continue
        q = ctx.Queue()
        port = 33000 + (seed % 2000)
        procs = [ctx.Process(target=worker, args=(r, world, port, seed, q)) for r in range(world)]
        for p in procs:
            p.start()
        res = {}
        deadline = time.time() + 240
        while len(res) < world and time.time() < deadline:
            try:
                r, blob = q.get(timeout=2)
                res[r] = pickle.loads(blob)
            except Exception:
                if any((not p.is_alive()) and p.exitcode for p in procs):
                    break
        for p in procs:
            p.join(timeout=20)
            if p.is_alive():
                p.kill()
        if len(res) < world:
            bad_cases += 1
            print(name, "A RANK DIED OR HUNG", case[1:])
            continue
        for r in range(world):
            got = res[r]
            if isinstance(got, str):
                bad_cases += 1
                print(name, "rank %d failed: %s" % (r, got[:200]), la)
                break
            for f in want:      # the unsharded run carries our own dict layout: give it the reference's sentinels
                f["mvs"][~f["searched"]] = 0
                f["mvs"][~f["searched"], 0, 0] = 0x7FFF
            if r == 0:
                bad = compare.compare_runs(want, got, check_planes=False, cutree=la.get("cuTree", 1), weightp=la.get("weightp", 1) or la.get("weightb", 0))
            else:
                # rank 0's output is the product; the other ranks run neither cuTree nor the estimates only rate control asks for
                # (DESIGN 6), but they must take the same decisions in the same order
                bad = ["decisions differ: %s / %s" % (a, b) for a, b in
                       zip([(f["poc"], f["sliceType"], f["bScenecut"]) for f in want], [(f["poc"], f["sliceType"], f["bScenecut"]) for f in got]) if a != b]
                if len(want) != len(got):
                    bad.append("frame count %d / %d" % (len(want), len(got)))
            if bad:
                bad_cases += 1
                print(name, "rank %d MISMATCH" % r, depth, w, h, n, skw, la)
                print("   ", "\n    ".join(bad[:4]))
                break
    print("%d cases over %d ranks: %d bad, %d refused by the host library, %.0f s" % (n_cases, world, bad_cases, refused, time.time() - t0))
    return 1 if bad_cases else 0


if __name__ == "__main__":
    sys.exit(main())
