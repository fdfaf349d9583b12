#!/bin/bash
# development aid: A/B of the even split of single-wave launches (K1s) -> gpurun_out/even_ab.txt
out=gpurun_out/even_ab.txt
: > $out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_even.log 2>&1; echo "pytest rc $?" >> $out; tail -2 gpurun_out/pytest_even.log >> $out
for rep in 1 2; do
  echo "== even" >> $out
  python tools/gpu_perf.py --cfgs cfg2,cfg3 --variants fast --iters 200 --graph >> $out 2>&1
  python tools/gpu_perf.py --cfgs cfg4 --p4 8192 --variants fast --iters 200 --graph >> $out 2>&1
  python tools/gpu_perf.py --cfgs cfg4 --p4 16384 --variants fast --iters 100 --graph >> $out 2>&1
  echo "== tiles of G*R" >> $out
  NTG_B200_NO_EVEN_SPLIT=1 python tools/gpu_perf.py --cfgs cfg2,cfg3 --variants fast --iters 200 --graph >> $out 2>&1
  NTG_B200_NO_EVEN_SPLIT=1 python tools/gpu_perf.py --cfgs cfg4 --p4 8192 --variants fast --iters 200 --graph >> $out 2>&1
  NTG_B200_NO_EVEN_SPLIT=1 python tools/gpu_perf.py --cfgs cfg4 --p4 16384 --variants fast --iters 100 --graph >> $out 2>&1
done
python tools/gpu_perf.py --cfgs cfg4 --variants fast,exact --iters 50 --graph >> $out 2>&1
cat $out
