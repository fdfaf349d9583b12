#!/bin/bash
# development aid: time every pack variant under build/variants/ (tools/gpu_perf.py, cfg5 fast)
for d in build/variants/*/; do
  t=$(basename $d)
  echo -n "$t: "
  NTG_B200_PACK_DIR=$d timeout 120 python tools/gpu_perf.py --cfgs ${CFG:-cfg5} --variants fast --iters 20 2>&1 | tail -1
done
