#!/bin/bash
# Forward build kernel: target-patch geometry A/B (profiles/fwd_store_pattern_r2.txt).  Builds variant libraries with the patch
# shape overridden at compile time into scratch_geo/ (git-ignored), then times pcfa_corr_pyramid_forward with each of them
# (PCFA_LIB selects the library; CUDA events, L2 flushed).  Run the build part where nvcc is, the timing part on a B200.
set -e
mkdir -p scratch_geo/obj
build_variant() {   # tag, extra nvcc flags
  for f in pcfa_b200/csrc/*.cu; do
    nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -Iinclude $2 \
         -c "$f" -o "scratch_geo/obj/$(basename "$f" .cu).o"
  done
  nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -shared -o "scratch_geo/libpcfa_$1.so" scratch_geo/obj/*.o
}
if [ "$1" = build ]; then
  build_variant g4x32 "-DPCFA_TC_PH=4 -DPCFA_TC_PW=32"     # shipped shape
  build_variant g8x16 "-DPCFA_TC_PH=8 -DPCFA_TC_PW=16"     # round-1 shape (in-warp level-1 pooling)
  exit 0
fi
for lib in g4x32 g8x16; do for f in 1 0; do
  echo "== $lib fuse_l1=$f"; PCFA_LIB=$PWD/scratch_geo/libpcfa_$lib.so PCFA_FWD_FUSE_L1=$f python scripts/bench_kernels.py 1 0 2>&1 | grep pyramid_forward
done; done
