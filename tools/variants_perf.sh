#!/bin/bash
# development aid: time every pack variant under build/variants/ (built by tools/variant_build.py: libntg_b200.so is
# found through their rpath, do not copy packs there by hand) -> gpurun_out/variants.txt; both steady-state
# instantiations of K1s (plain / with the push epilogue), the dense layout, and two mid-size batches
out=gpurun_out/variants.txt
: > $out
for rep in 1 2; do
for d in build/variants/*/; do
  t=$(basename $d)
  echo -n "$t plain: " >> $out
  NTG_B200_NO_PUSH_KERNEL=1 NTG_B200_PACK_DIR=$d timeout 120 python tools/gpu_perf.py --cfgs ${CFG:-cfg4} --variants fast --iters 50 --graph 2>&1 | tail -1 >> $out
  echo -n "$t push: " >> $out
  NTG_B200_PACK_DIR=$d timeout 120 python tools/gpu_perf.py --cfgs ${CFG:-cfg4} --variants fast --iters 50 --graph 2>&1 | tail -1 >> $out
  echo -n "$t dense: " >> $out
  NTG_B200_PACK_DIR=$d timeout 120 python tools/gpu_perf.py --cfgs ${CFG:-cfg4} --variants fast --iters 50 --graph --dense 2>&1 | tail -1 >> $out
  echo -n "$t 32768: " >> $out
  NTG_B200_PACK_DIR=$d timeout 120 python tools/gpu_perf.py --cfgs ${CFG:-cfg4} --variants fast --iters 50 --graph --p4 32768 2>&1 | tail -1 >> $out
  echo -n "$t 8192: " >> $out
  NTG_B200_PACK_DIR=$d timeout 120 python tools/gpu_perf.py --cfgs ${CFG:-cfg4} --variants fast --iters 100 --graph --p4 8192 2>&1 | tail -1 >> $out
done
done
cat $out
