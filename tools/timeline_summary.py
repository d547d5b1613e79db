"""Summarises an X265CU_TIMELINE dump (engine.cu, resolveProfile): for the last complete step, per kernel family the busy
union, and a coarse text timeline (one row per family, one column per time bin: share of the bin the family was running)."""
import sys
import numpy as np

NAMES = ["lowres", "aq", "intra", "search", "cost", "weight", "cutree"]


def main(path, bins=100):
    steps, cur = [], []
    for line in open(path):
        k, a, b = line.strip().split(",")
        if int(k) < 0:
            if cur:
                steps.append(cur)
            cur = []
        else:
            cur.append((int(k), float(a), float(b)))
    if not steps:
        print("no complete step")
        return
    which = int(sys.argv[3]) if len(sys.argv) > 3 else -1
    st = steps[which]
    print("step %d of %d" % (which % len(steps), len(steps)))
    t0 = min(a for _, a, _ in st); t1 = max(b for _, _, b in st)
    print("step of %d launch brackets, %.1f ms" % (len(st), t1 - t0))
    edges = np.linspace(t0, t1, bins + 1)
    anyb = np.zeros(bins)
    for k, name in enumerate(NAMES):
        iv = sorted((a, b) for kk, a, b in st if kk == k)
        occ = np.zeros(bins)
        end = -1e30
        busy = 0.0
        for a, b in iv:
            a2 = max(a, end)
            if b > a2:
                busy += b - a2
                lo, hi = np.searchsorted(edges, [a2, b])
                for i in range(max(lo - 1, 0), min(hi, bins)):
                    occ[i] += max(0.0, min(b, edges[i + 1]) - max(a2, edges[i]))
                end = b
        occ /= (edges[1] - edges[0])
        anyb = np.maximum(anyb, occ)
        row = "".join(" .:-=+*#%@"[min(9, int(o * 9.999))] for o in occ)
        print("%-7s busy %7.1f ms |%s|" % (name, busy, row))
    print("%-7s               |%s|" % ("any", "".join(" .:-=+*#%@"[min(9, int(o * 9.999))] for o in anyb)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 100)
