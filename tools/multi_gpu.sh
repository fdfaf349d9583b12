#!/bin/bash
# round-end evidence on N GPUs: the driver's own command -> gpurun_out/r02_final_bench_cfg4_${N}gpu.json
N=${1:-2}
if [ "$2" = "pytest" ]; then python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final_$N.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/pytest_final_$N.log; fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 50 --warmup 5 2> gpurun_out/multi_err_$N.log | tail -1 > gpurun_out/r02_final_bench_cfg4_${N}gpu.json
python tools/show_bench.py gpurun_out/r02_final_bench_cfg4_${N}gpu.json || tail -5 gpurun_out/multi_err_$N.log
