#!/bin/bash
out=gpurun_out/e2e_ab.txt
: > $out
for rep in 1 2; do
for mb in 32 64 128 192 256 400; do NTG_B200_HOST_CHUNK_MB=$mb python tools/gpu_e2e.py 2>&1 | tail -1 >> $out; done
done
cat $out
