#!/bin/bash
out=gpurun_out/even2.txt
: > $out
for rep in 1 2; do
for p in 24576 49152 65536; do
  echo "== P=$p default" >> $out
  python tools/gpu_perf.py --cfgs cfg4 --p4 $p --variants fast --iters 100 --graph >> $out 2>&1
  echo "== P=$p even forced (ties, up to 32 tiles per CTA)" >> $out
  NTG_B200_EVEN_MAXTILES=32 NTG_B200_EVEN_TIE=1 python tools/gpu_perf.py --cfgs cfg4 --p4 $p --variants fast --iters 100 --graph >> $out 2>&1
done
done
NTG_B200_DEBUG=1 NTG_B200_EVEN_MAXTILES=32 NTG_B200_EVEN_TIE=1 python tools/gpu_perf.py --cfgs cfg4 --variants fast --iters 1 2>&1 | grep "K1s launch" | tail -1 >> $out
cat $out
