"""Regenerates tests/golden/*.npz from the unmodified reference (oracle/_ref, built from
/root/reference by oracle/Makefile.ref).  Run in the build container: python tests/golden/make_golden.py [names...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import _pkg
import cases
import golden_io
import refbind

if __name__ == "__main__":
    synth = _pkg.load_synth()
    for name in (sys.argv[1:] or cases.GOLDEN):     # optionally only the named fixtures
        case = cases.get_case(name)
        frames = cases.run_reference(refbind, synth, case, planes=True)
        golden_io.save(name, frames)
        print(name, len(frames), os.path.getsize(golden_io.path(name)))
