#!/bin/bash
# leaf size x walk-loop thresholds of k_pt_warp_bvh on config 4 (1/16 frame) and config 5 (full frame, fp64 + fp32)
V="python scratch/variants.py --iters 4 --warm 1"
for lib in t4 t2 t1; do
  echo "== leaf variant $lib"
  $V --workload c4 --scale 4 --accel bvh scratch/v/$lib.so
  $V --workload c5 --accel bvh scratch/v/$lib.so
  $V --workload c5 --accel bvh --precision f32 scratch/v/$lib.so
done
for lib in t4 t2; do
for ra in 12 16 24 28; do for im in 4 8 16; do
  echo "== $lib refill_at $ra inner_min $im"
  RT_BVH_REFILL_AT=$ra RT_BVH_INNER_MIN=$im $V --workload c4 --scale 4 --accel bvh scratch/v/$lib.so
done; done; done
