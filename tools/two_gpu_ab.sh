#!/bin/bash
# development aid: old tree (build/old_tree) against the current one on N GPUs -> gpurun_out/two_gpu_ab.txt
N=${1:-2}
out=$PWD/gpurun_out/two_gpu_ab.txt
: > $out
run() { # dir, env, args
  echo "== $1 $2 $3" >> $out
  (cd $1 && env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus $N --steps 50 --warmup 5 --no-others $3 2> /tmp/err.log | tail -1 > /tmp/line.json)
  python tools/show_bench.py /tmp/line.json 2>&1 | head -1 >> $out || tail -5 /tmp/err.log >> $out
}
for rep in 1 2; do
run . A=1 ""
run build/old_tree A=1 ""
run . NTG_B200_NO_PUSH_ROTATE=1 ""
run . A=1 "--scaling strong"
run build/old_tree A=1 "--scaling strong"
run . NTG_B200_NO_EVEN_SPLIT=1 "--scaling strong"
done
cat $out
