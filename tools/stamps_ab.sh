#!/bin/bash
out=gpurun_out/stamps_ab.txt
: > $out
for v in stamps stamps_late; do
echo "== $v even" >> $out
NTG_STAMPS_VARIANT=$v python tools/gpu_stamps.py >> $out 2>&1
echo "== $v tiles of G*R" >> $out
NTG_STAMPS_VARIANT=$v NTG_B200_NO_EVEN_SPLIT=1 python tools/gpu_stamps.py >> $out 2>&1
done
grep -v "launch 2[012]" $out
