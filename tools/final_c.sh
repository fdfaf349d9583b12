#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final_c.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/pytest_final_c.log
python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc $?"
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/bench_final_err.log; echo "bench rc $?"
python tools/gpu_perf.py --cfgs cfg4 --variants fast --iters 3 > gpurun_out/perf4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ntg_eval_small -s 3 -c 1 -f -o gpurun_out/r02_k1s_cfg4_final python tools/gpu_perf.py --cfgs cfg4 --variants fast --iters 3 > gpurun_out/ncu4.log 2>&1
python tools/show_bench.py gpurun_out/r02_bench_final.json
python tools/gpu_perf.py --cfgs cfg2,cfg3 --variants fast --iters 200 --graph
NTG_B200_NO_PUSH_KERNEL=1 python tools/gpu_perf.py --cfgs cfg2,cfg3,cfg4 --variants fast --iters 100 --graph
