#!/bin/bash
out=gpurun_out/dense_ab.txt
: > $out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_dense.log 2>&1; echo "pytest rc $?" >> $out; tail -2 gpurun_out/pytest_dense.log >> $out
for rep in 1 2; do
python tools/gpu_perf.py --cfgs cfg4 --variants fast,exact --iters 50 --graph --dense >> $out 2>&1
done
python tools/gpu_perf.py --cfgs cfg2,cfg3 --variants fast --iters 200 --graph --dense >> $out 2>&1
python tools/gpu_perf.py --cfgs cfg4 --variants fast --iters 50 --graph >> $out 2>&1
cat $out
