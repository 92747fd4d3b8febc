#!/bin/bash
# one GPU: full GPU suite; true cold start (no cubin disk cache); launch list; L2 fetch granularity; solver phases; other decks
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/g2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g2_pytest.log
tail -8 gpurun_out/g2_pytest.log
export RUN_TIMEOUT=600
# cold start without the cubin disk cache, in a fresh process: NVRTC in the background, interpreted fill when sooner
SPED_CACHE_DIR= SPED_LOG=1 tools/run_n.sh 1 heisenberg_square_6x6 g2_cold_nocache --steps 20 --no-cpu --no-parity --no-sharded-at-one
grep -E "jit|cold eigh|operator cache" gpurun_out/g2_cold_nocache.err | head -12
SPED_CACHE_DIR= SPED_JIT_WAIT=1 SPED_LOG=1 tools/run_n.sh 1 heisenberg_square_6x6 g2_cold_nocache_wait --steps 20 --no-cpu --no-parity --no-sharded-at-one
grep -E "jit|cold eigh|operator cache" gpurun_out/g2_cold_nocache_wait.err | head -12
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02b_launches_6x6.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-parity --no-eigh --no-sharded-at-one > gpurun_out/g2_ncu_list.log 2>&1
for f in 32 64 128; do
  SPED_L2_FETCH=$f tools/run_n.sh 1 heisenberg_chain_36 g2_c36_fetch$f --steps 20 --no-cpu --no-parity --no-eigh --no-sharded-at-one --e2e-host-gb 0
  SPED_L2_FETCH=$f tools/run_n.sh 1 heisenberg_square_6x6 g2_6x6_fetch$f --steps 20 --no-cpu --no-parity --no-eigh --no-sharded-at-one --e2e-host-gb 0
done
tools/run_n.sh 1 heisenberg_chain_36 g2_c36_default --steps 20 --no-cpu --no-parity --no-eigh --no-sharded-at-one --e2e-host-gb 0
for d in heisenberg_square_6x6 xxz_triangular_19 heisenberg_chain_24 heisenberg_pyrochlore_32 heisenberg_triangular_19; do
  timeout 300 python tools/eigh_probe.py $d 2>&1 | grep -E "first|second|dropped:"
done
for d in xxz_triangular_19 heisenberg_pyrochlore_32 heisenberg_chain_24 heisenberg_triangular_19; do tools/run_n.sh 1 $d g2_$d --steps 50 --cpu-seconds 5 --no-sharded-at-one; done
