#!/bin/bash
mkdir -p gpurun_out
export RUN_TIMEOUT=300
for v in 1 15 17 19 21 23 25; do
  SPED_CACHED_VARIANT=$v tools/run_n.sh 1 heisenberg_square_6x6 v4_var$v --steps 30 --no-cpu --no-eigh --no-parity --e2e-host-gb 0
done
for v in 1 15 19 25; do
  SPED_CACHED_VARIANT=$v tools/run_n.sh 1 heisenberg_chain_36 v4_c36_var$v --steps 20 --no-cpu --no-eigh --no-parity --e2e-host-gb 0
done
