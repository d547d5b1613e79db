"""A fixed slice of the randomised sweep (tools/fuzz_host_vs_reference.py) in the CPU suite: the seeds that found bugs in round 2
plus a few dozen others.  Needs the live reference (oracle/_ref)."""
import os
import subprocess
import sys

import pytest

import refbind

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# 150 / 254 / 443 / 476 / 723: a speculated pair whose reference had been released (rc-lookahead barely above bframes);
# 1219: the 12-bit mv-cost table; 20 / 21 / 30 / 31 / 35: aq-mode 0 + weightp; 81 / 147: hist-scenecut on 4:0:0 (the reference dies)
SEEDS = [150, 254, 443, 476, 723, 1219, 20, 21, 30, 31, 35, 81, 147] + list(range(0, 20)) + list(range(9000, 9012))


def test_fixed_slice_of_the_randomised_sweep():
    if not all(refbind.available(d) for d in (8, 10, 12)):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_host_vs_reference.py"), "seeds", ",".join(str(s) for s in SEEDS)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    tail = r.stdout.strip().splitlines()[-1]
    assert r.returncode == 0, r.stdout[-3000:]
    assert "mismatch 0" in tail and "ours_crashed 0" in tail and "ours_refused 0" in tail, tail
