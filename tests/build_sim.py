"""Builds the CPU test doubles under tests/_build/ (TEST INFRASTRUCTURE):
  liboracle{8,10}.so      -- oracle/la_oracle.c, the C restatement
  libx265la_sim{8,10}.so  -- the PRODUCT host logic (x265-amod_b200/host/*.cpp) linked against
                             tests/simengine/simengine.cpp (the engine ABI on top of the oracle)
Nothing here is shipped or loaded by the product."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "_build")
HOST = os.path.join(ROOT, "x265-amod_b200", "host")


def _run(cmd):
    subprocess.run(cmd, check=True)


def _stale(target, srcs):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force=False):
    os.makedirs(OUT, exist_ok=True)
    oc = os.path.join(ROOT, "oracle", "la_oracle.c")
    oh = os.path.join(ROOT, "oracle", "la_oracle.h")
    host_srcs = [os.path.join(HOST, f) for f in ("lookahead.cpp", "la_capi.cpp")]
    host_hdrs = [os.path.join(HOST, f) for f in ("lookahead.h", "la_capi.h")] + [os.path.join(ROOT, "include", "x265cu.h"),
                 os.path.join(ROOT, "x265-amod_b200", "csrc", "la_me_generic.cuh")]
    sim = os.path.join(ROOT, "tests", "simengine", "simengine.cpp")
    for d in (8, 10, 12):
        lib = os.path.join(OUT, "liboracle%d.so" % d)
        if force or _stale(lib, [oc, oh]):
            _run(["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-DOR_DEPTH=%d" % d, "-o", lib, oc, "-lm"])
        obj = os.path.join(OUT, "la_oracle%d.o" % d)
        lib = os.path.join(OUT, "libx265la_sim%d.so" % d)
        if force or _stale(lib, [oc, oh, sim] + host_srcs + host_hdrs):
            _run(["gcc", "-O2", "-std=c99", "-fPIC", "-c", "-DOR_DEPTH=%d" % d, "-o", obj, oc])
            _run(["g++", "-O2", "-std=c++11", "-Wno-unknown-pragmas", "-fPIC", "-shared", "-DOR_DEPTH=%d" % d,
                  "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle"), "-I" + HOST,
                  "-o", lib, sim, obj] + host_srcs + ["-lm"])
    return OUT


if __name__ == "__main__":
    build(force=True)
    print("built", OUT)
