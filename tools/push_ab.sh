#!/bin/bash
out=$PWD/gpurun_out/push_ab.txt
: > $out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_p.log 2>&1; echo "pytest rc $?" >> $out; tail -2 gpurun_out/pytest_p.log >> $out
NTG_B200_FORCE_PEERS_KERNEL=1 python -m pytest tests -m gpu -x -q -k "even_split or ragged or baseline_sizes" > gpurun_out/pytest_p2.log 2>&1; echo "pytest (push kernel) rc $?" >> $out; tail -2 gpurun_out/pytest_p2.log >> $out
for rep in 1 2; do
for d in . build/old_tree; do
  echo "== $d force-peers kernel" >> $out
  (cd $d && NTG_B200_FORCE_PEERS_KERNEL=1 python tools/gpu_perf.py --cfgs cfg4 --variants fast --iters 50 --graph) >> $out 2>&1
  echo "== $d" >> $out
  (cd $d && python tools/gpu_perf.py --cfgs cfg4 --variants fast,exact --iters 50 --graph) >> $out 2>&1
done
done
python tools/gpu_perf.py --cfgs cfg2,cfg3 --variants fast --iters 200 --graph >> $out 2>&1
for p in 4096 8192 16384 32768; do python tools/gpu_perf.py --cfgs cfg4 --p4 $p --variants fast --iters 100 --graph >> $out 2>&1; done
python tools/gpu_perf.py --cfgs cfg4 --variants fast --iters 50 --graph --dense >> $out 2>&1
cat $out
