#!/bin/bash
# final validation of the committed state: full GPU suite, smoke(), bench N=1 (both arms quick)
mkdir -p gpurun_out
( time timeout -s KILL 2400 python -m pytest tests -m gpu -q 2>&1 | tail -5 ) 2>&1 | tee gpurun_out/r51_tests.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -s KILL 900 python bench.py 2>gpurun_out/r51_bench.err > gpurun_out/r51_bench.json; cut -c1-260 gpurun_out/r51_bench.json
