#!/bin/bash
# full GPU suite (without -x so that every failure shows) on the state with pcm.cu / wav_io / feature_extraction
mkdir -p gpurun_out
( time timeout -s KILL 3000 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) 2>&1 | tee gpurun_out/r36_tests.log
