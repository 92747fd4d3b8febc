#!/bin/bash
mkdir -p gpurun_out
export RUN_TIMEOUT=300
for v in 1 3 5 7 9 11 13; do
  SPED_CACHED_VARIANT=$v tools/run_n.sh 1 heisenberg_square_6x6 v3_var$v --steps 30 --no-cpu --no-eigh --e2e-host-gb 0
done
for v in 1 3 9; do
  SPED_CACHED_VARIANT=$v tools/run_n.sh 1 heisenberg_chain_36 v3_c36_var$v --steps 20 --no-cpu --no-eigh --e2e-host-gb 0
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cached_matvec_kernel -s 4 -c 1 -o gpurun_out/r02a_prof_cached_6x6 python bench.py --steps 3 --warmup 3 --no-eigh --no-cpu --e2e-host-gb 0 > gpurun_out/v3_ncu.log 2>&1
tail -2 gpurun_out/v3_ncu.log
