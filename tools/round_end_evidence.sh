#!/bin/bash
# round-end evidence: the bench line, the launch list of the same command, one full ncu capture per roofline kernel
set -x
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/bench_final_err.log; echo "bench rc $?"
python bench.py --no-others --steps 20 --warmup 3 > gpurun_out/bench_small.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --no-others --steps 20 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
python tools/gpu_perf.py --cfgs cfg4 --variants fast --iters 3 > gpurun_out/perf4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ntg_eval_small -s 3 -c 1 -f -o gpurun_out/r02_k1s_cfg4_final python tools/gpu_perf.py --cfgs cfg4 --variants fast --iters 3 > gpurun_out/ncu4.log 2>&1
python tools/gpu_perf.py --cfgs cfg5 --p5 16384 --variants fast --iters 2 > gpurun_out/perf5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cluster_hot -s 3 -c 1 -f -o gpurun_out/r02_k1ch_cfg5_final python tools/gpu_perf.py --cfgs cfg5 --p5 16384 --variants fast --iters 2 > gpurun_out/ncu5.log 2>&1
ls -la gpurun_out/*.ncu-rep
python tools/show_bench.py gpurun_out/r02_bench_final.json
