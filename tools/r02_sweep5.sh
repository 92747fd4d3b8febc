#!/bin/bash
mkdir -p gpurun_out
export RUN_TIMEOUT=300
for v in 19 27 29 31 33 35; do
  SPED_CACHED_VARIANT=$v tools/run_n.sh 1 heisenberg_square_6x6 v5_var$v --steps 30 --no-cpu --no-eigh --no-parity --e2e-host-gb 0
done
for v in 27 29 33; do
  SPED_CACHED_VARIANT=$v tools/run_n.sh 1 heisenberg_chain_36 v5_c36_var$v --steps 20 --no-cpu --no-eigh --no-parity --e2e-host-gb 0
done
