#!/bin/bash
# development aid: 2-GPU checks of the fused gather with the current kernels -> gpurun_out/two_gpu.txt
N=${1:-2}
out=gpurun_out/two_gpu.txt
: > $out
run() {
  echo "== $*" >> $out
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus $N --steps 50 --warmup 5 --no-others "$@" 2> gpurun_out/two_gpu_err.log | tail -1 > gpurun_out/two_gpu_line.json
  python tools/show_bench.py gpurun_out/two_gpu_line.json >> $out 2>&1 || tail -5 gpurun_out/two_gpu_err.log >> $out
}
run
cp gpurun_out/two_gpu_line.json gpurun_out/bench_${N}gpu_weak.json
run --scaling strong
run --scaling strong --problems 16384
run --scaling strong --problems 8192
python -m pytest tests -m gpu -x -q -k "fused_peer" >> $out 2>&1
cat $out
