#!/bin/bash
out=gpurun_out/final_a.txt
: > $out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc $?" >> $out; tail -2 gpurun_out/pytest_final.log >> $out
NTG_B200_FORCE_PEERS_KERNEL=1 python -m pytest tests -m gpu -x -q -k "even_split or ragged or baseline_sizes" > gpurun_out/pytest_final2.log 2>&1; echo "pytest (push kernel) rc $?" >> $out; tail -2 gpurun_out/pytest_final2.log >> $out
python -c "import __graft_entry__ as g; g.smoke()" >> $out 2>&1; echo "smoke rc $?" >> $out
python tools/gpu_perf.py --cfgs cfg2,cfg3 --variants fast --iters 200 --graph >> $out 2>&1
for p in 4096 8192 16384 32768; do python tools/gpu_perf.py --cfgs cfg4 --p4 $p --variants fast --iters 100 --graph >> $out 2>&1; done
python tools/gpu_perf.py --cfgs cfg4 --variants fast,exact --iters 50 --graph >> $out 2>&1
python tools/gpu_perf.py --cfgs cfg4 --variants fast,exact --iters 50 --graph --dense >> $out 2>&1
python tools/gpu_perf.py --cfgs cfg5 --p5 16384 --variants fast --iters 5 >> $out 2>&1
cat $out
