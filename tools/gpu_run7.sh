#!/bin/bash
mkdir -p gpurun_out
for f in 12 11 10 9 8 7 6 5 4 3; do MST_ENC_FIRST_UMMA=$f timeout 300 python tools/enc_debug.py 2 32768 2>&1 | tail -1; done | tee gpurun_out/encdbg.log
for f in 12 9 3; do MST_ENC_FIRST_UMMA=$f timeout 300 python tools/enc_debug.py 1 262144 2>&1 | tail -1; done | tee -a gpurun_out/encdbg.log
