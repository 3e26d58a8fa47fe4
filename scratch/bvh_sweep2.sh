#!/bin/bash
V="python scratch/variants.py --iters 4 --warm 1 --workload c4 --scale 4 --accel bvh"
$V scratch/v/b2.so scratch/v/b3.so
for ra in 8 12 16 20 24; do for im in 4 8 12; do
  echo "== refill_at $ra inner_min $im"
  RT_BVH_REFILL_AT=$ra RT_BVH_INNER_MIN=$im $V scratch/v/b2t.so
done; done
