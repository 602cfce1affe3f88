#!/bin/bash
# tries the launch shapes of k_count_map_q (HLAC_COUNT_SHAPE=smem_bitmaps,threads,slots_per_warp,idx_in_smem; HLAC_COUNT_TMA=1 =
# bulk-copy record prefetch) on config 2's read mix; results: profiles/r02_count_map_tuning.md
N=${1:-10000000}
for s in 1,1024,64,1 1,1024,48,1 1,1024,56,0 1,1024,64,0 1,1024,40,0 1,512,64,0 0,256,64,0 0,128,64,0; do
  for t in 0 1; do
    echo -n "shape $s tma $t  "; HLAC_COUNT_TMA=$t HLAC_COUNT_SHAPE=$s timeout 90 python tools/prof_step.py $N 6 2>&1 | tail -1 | sed -e 's/.*| \(smem_bitmaps.*\)/\1/'
  done
done
