#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python bench.py --no-cpu-baseline 2>gpurun_out/r55.err | tail -1 > gpurun_out/r55_bench.json
python -c "
import json
d=json.load(open('gpurun_out/r55_bench.json'))
print(d['value'], d['ms_per_step']); print(d['extra']['wav_io']); print(d['extra']['fx_chain_config3']['ms'])"
tail -3 gpurun_out/r55.err
