#!/bin/bash
# development aid: A/B of the per-CTA rotation of the peer-store order (K1s PUSH epilogue) on N GPUs
# usage (on the GPU box): tools/rot_ab.sh N  -> gpurun_out/rot_ab.txt
N=${1:-8}
out=gpurun_out/rot_ab.txt
: > $out
run() { # label, env assignment, extra args
  echo "== $1 $3" >> $out
  env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus $N --steps 50 --warmup 5 --no-others $3 2> gpurun_out/rot_ab_err.log | tail -1 > gpurun_out/rot_ab_line.json
  python tools/show_bench.py gpurun_out/rot_ab_line.json >> $out 2>&1 || tail -5 gpurun_out/rot_ab_err.log >> $out
}
for rep in 1 2; do
  run rotate NTG_B200_DUMMY=1 ""
  run norot NTG_B200_NO_PUSH_ROTATE=1 ""
  run rotate NTG_B200_DUMMY=1 "--scaling strong"
  run norot NTG_B200_NO_PUSH_ROTATE=1 "--scaling strong"
done
cat $out
