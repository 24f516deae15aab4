#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:tcn_block_umma -s 21 -c 2 -o gpurun_out/tcn_r53 -f python tools/tcn_time.py > gpurun_out/r53_ncu.log 2>&1
tail -2 gpurun_out/r53_ncu.log; ls -la gpurun_out/tcn_r53.ncu-rep
