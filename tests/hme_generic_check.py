"""Helper of test_generic_search_source_on_cpu: runs one --hme case through the CPU sim engine with X265SIM_GENERIC_ME=1, i.e.
with the per-block searches done by the PRODUCT's la_me_generic.cuh compiled for the CPU (tests/simengine/simengine.cpp), and
compares every published field with the live reference.  Prints the mismatches; exit code 0 = none."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)


def main(name):
    assert os.environ.get("X265SIM_GENERIC_ME") == "1"
    import _pkg
    import build_sim
    import cases
    import compare
    import refbind
    pkg = _pkg.load_pkg(); synth = _pkg.load_synth()
    case = cases.get_case(name)
    simdir = build_sim.build()
    want = cases.run_reference(refbind, synth, case)
    got = cases.run_ours(pkg, synth, case, lib_path=os.path.join(simdir, "libx265la_sim%d.so" % case[1]))
    bad = compare.compare_runs(want, got, check_planes=True, weightp=case[6].get("weightp", 1))
    print("\n".join(bad[:10]))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1]))
