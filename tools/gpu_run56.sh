#!/bin/bash
mkdir -p gpurun_out
( time timeout -s KILL 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/r56_tests.log
timeout -s KILL 400 python bench.py --no-cpu-baseline 2>gpurun_out/r56.err | tail -1 > gpurun_out/r56_bench.json
python -c "
import json
d=json.load(open('gpurun_out/r56_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']); print(d['extra']['wav_io'])"
