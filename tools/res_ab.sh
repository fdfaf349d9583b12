#!/bin/bash
python tools/gpu_e2e.py resident
NTG_B200_NO_PUSH_KERNEL=1 python tools/gpu_e2e.py resident
python tools/gpu_e2e.py resident
