#!/bin/bash
# streaming kernel after the rewrite (no window, default class as its own loop): variants + block kernel
mkdir -p gpurun_out
export RUN_TIMEOUT=300
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/v2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v2_pytest.log
tail -3 gpurun_out/v2_pytest.log
tools/run_n.sh 1 heisenberg_square_6x6 v2_default --steps 30 --no-cpu
SPED_CACHED_VARIANT=1 tools/run_n.sh 1 heisenberg_square_6x6 v2_var1 --steps 30 --no-cpu --no-eigh
SPED_CACHED_VARIANT=2 tools/run_n.sh 1 heisenberg_square_6x6 v2_var2 --steps 30 --no-cpu --no-eigh
SPED_DEFAULT_CLASS=0 tools/run_n.sh 1 heisenberg_square_6x6 v2_nodefault --steps 30 --no-cpu --no-eigh
SPED_CACHED_BLOCKS_PER_SM=4 tools/run_n.sh 1 heisenberg_square_6x6 v2_4blocks --steps 30 --no-cpu --no-eigh
timeout 300 python tools/block_bench.py heisenberg_square_6x6 > gpurun_out/v2_block_6x6.json 2> gpurun_out/v2_block_6x6.err; cat gpurun_out/v2_block_6x6.json
timeout 300 python tools/block_bench.py xxz_triangular_19 50 > gpurun_out/v2_block_xxz.json 2> gpurun_out/v2_block_xxz.err; cat gpurun_out/v2_block_xxz.json
tools/run_n.sh 1 heisenberg_chain_36 v2_chain36 --steps 20 --no-cpu --no-eigh --e2e-host-gb 0
