# kernel tuning variants of libx265cu.so (+ the host library beside each): tools/build_variants.sh "name:-DFLAG=1 ..." ...
# builds x265-amod_b200/lib_<name>/; run one with X265CU_LIBDIR=$PWD/x265-amod_b200/lib_<name> python bench.py ...
set -e
cd "$(dirname "$0")/.."
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  d=x265-amod_b200/lib_$name; mkdir -p $d
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++14 -Xcompiler -fPIC -shared -Iinclude -Ix265-amod_b200/csrc $flags -o $d/libx265cu.so x265-amod_b200/csrc/engine.cu
  g++ -O2 -std=c++11 -fPIC -shared -Iinclude -Ix265-amod_b200/host -o $d/libx265la.so x265-amod_b200/host/lookahead.cpp x265-amod_b200/host/la_capi.cpp -L$d -lx265cu '-Wl,-rpath,$ORIGIN'
  echo "built $d ($flags)"
done
