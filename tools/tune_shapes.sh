#!/bin/bash
# tries the launch shapes of k_count_map (HLAC_COUNT_SHAPE=smem_bitmaps,R,threads[,min_blocks]) on 4 M reads;
# results of the r01 sweep: profiles/r01_count_map_tuning.md
for s in 1,2,1024 1,1,512 1,1,256 1,4,512 1,2,512 0,2,128 0,1,128 0,2,256 0,1,256 0,4,128; do
  echo -n "$s  "; HLAC_COUNT_SHAPE=$s python tools/prof_step.py 4000000 4 2>&1 | tail -1 | sed -e 's/.*map_ms.: \([0-9.]*\).*/\1/'
done
