#!/bin/bash
# forward build kernel: target-patch geometry (8x16 vs 4x32) x level-1 pooling in the epilogue (on/off); events, L2 flushed
for lib in g8x16 g4x32; do for f in 1 0; do
  [ "$lib" = g4x32 ] && [ "$f" = 1 ] && continue     # the fused pooling assumes 8x16 patches
  echo "== $lib fuse_l1=$f"; PCFA_LIB=$PWD/scratch_geo/libpcfa_$lib.so PCFA_FWD_FUSE_L1=$f python scripts/bench_kernels.py 1 0 2>&1 | grep pyramid_forward
done; done
