#!/bin/bash
# Round-2 validation of the recovered branch on ONE GPU: GPU test-suite, then 6x6 with each switch.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/v1_smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/v1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v1_pytest.log
tail -5 gpurun_out/v1_pytest.log
export RUN_TIMEOUT=400
tools/run_n.sh 1 heisenberg_square_6x6 v1_all_on --steps 30 --no-cpu
SPED_DEFAULT_CLASS=0 SPED_WINDOW=0 SPED_JIT_TOPALIGN=0 tools/run_n.sh 1 heisenberg_square_6x6 v1_all_off --steps 30 --no-cpu
SPED_WINDOW=0 tools/run_n.sh 1 heisenberg_square_6x6 v1_nowindow --steps 30 --no-cpu --no-eigh
SPED_DEFAULT_CLASS=0 tools/run_n.sh 1 heisenberg_square_6x6 v1_nodefault --steps 30 --no-cpu --no-eigh
SPED_WINDOW=0 SPED_CACHED_VARIANT=1 tools/run_n.sh 1 heisenberg_square_6x6 v1_nowindow_var1 --steps 30 --no-cpu --no-eigh
SPED_CACHED_VARIANT=1 tools/run_n.sh 1 heisenberg_square_6x6 v1_var1 --steps 30 --no-cpu --no-eigh
SPED_JIT_TOPALIGN=0 tools/run_n.sh 1 heisenberg_square_6x6 v1_notopalign --steps 30 --no-cpu --no-eigh
tools/run_n.sh 1 xxz_triangular_19 v1_xxz --steps 30 --no-cpu
tools/run_n.sh 1 heisenberg_pyrochlore_32 v1_pyro --steps 30 --no-cpu
tools/run_n.sh 1 heisenberg_chain_24 v1_chain24 --steps 30 --no-cpu
