#!/bin/bash
# new bench.py on one GPU (all legs incl. oracle parity + CPU baseline), GPU tests, block bench, reference arm
mkdir -p gpurun_out
nproc > gpurun_out/v4_nproc.txt; free -g | head -2 >> gpurun_out/v4_nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/v4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v4_pytest.log
tail -3 gpurun_out/v4_pytest.log
export RUN_TIMEOUT=600
tools/run_n.sh 1 heisenberg_square_6x6 v4_bench --steps 50
timeout 300 python tools/block_bench.py heisenberg_square_6x6 > gpurun_out/v4_block_6x6.json 2> gpurun_out/v4_block_6x6.err; cat gpurun_out/v4_block_6x6.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/v4_reference.json 2> gpurun_out/v4_reference.err; cut -c1-400 gpurun_out/v4_reference.json
