#!/bin/bash
out=gpurun_out/img_ab.txt
: > $out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_img.log 2>&1; echo "pytest rc $?" >> $out; tail -2 gpurun_out/pytest_img.log >> $out
NTG_B200_FORCE_PEERS_KERNEL=1 python -m pytest tests -m gpu -x -q -k "even_split or ragged or baseline_sizes" > gpurun_out/pytest_img2.log 2>&1; echo "pytest (push kernel) rc $?" >> $out; tail -2 gpurun_out/pytest_img2.log >> $out
for rep in 1 2; do
  echo "== even where the launcher takes it" >> $out
  python tools/gpu_perf.py --cfgs cfg2,cfg3 --variants fast --iters 200 --graph >> $out 2>&1
  for p in 4096 8192 16384; do python tools/gpu_perf.py --cfgs cfg4 --p4 $p --variants fast --iters 200 --graph >> $out 2>&1; done
  echo "== tiles of G*R" >> $out
  NTG_B200_NO_EVEN_SPLIT=1 python tools/gpu_perf.py --cfgs cfg2,cfg3 --variants fast --iters 200 --graph >> $out 2>&1
  for p in 4096 8192; do NTG_B200_NO_EVEN_SPLIT=1 python tools/gpu_perf.py --cfgs cfg4 --p4 $p --variants fast --iters 200 --graph >> $out 2>&1; done
done
python tools/gpu_perf.py --cfgs cfg4 --variants fast,exact --iters 50 --graph >> $out 2>&1
NTG_B200_DEBUG=1 python tools/gpu_perf.py --cfgs cfg4 --p4 8192 --variants fast --iters 1 2>&1 | grep "K1s launch" | tail -1 >> $out
echo "== stamps" >> $out
python tools/gpu_stamps.py 2>&1 | grep -v "launch 2[012]" >> $out
cat $out
