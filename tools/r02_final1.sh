#!/bin/bash
# one GPU: full GPU suite, sanitizers, the driver's bench line, ncu launch list + full capture of the dominant kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/f1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f1_pytest.log
tail -4 gpurun_out/f1_pytest.log
bash tools/sanitize.sh
export RUN_TIMEOUT=600
tools/run_n.sh 1 heisenberg_square_6x6 f1_bench --steps 50
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/f1_reference.json 2> gpurun_out/f1_reference.err; cut -c1-300 gpurun_out/f1_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_6x6.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-parity > gpurun_out/f1_ncu_list.log 2>&1
tail -1 gpurun_out/f1_ncu_list.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cached_matvec_kernel -s 4 -c 1 -o gpurun_out/r02_prof_cached_6x6 python bench.py --steps 3 --warmup 3 --no-eigh --no-cpu --no-parity --e2e-host-gb 0 > gpurun_out/f1_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cached_block_kernel -s 1 -c 1 -o gpurun_out/r02_prof_block_6x6 python tools/block_bench.py heisenberg_square_6x6 3 > gpurun_out/f1_ncu_block.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sped_cache_fill_jit -c 1 -o gpurun_out/r02_prof_fill_6x6 python bench.py --steps 3 --warmup 3 --no-eigh --no-cpu --no-parity --e2e-host-gb 0 > gpurun_out/f1_ncu_fill.log 2>&1
ls -la gpurun_out/*.ncu-rep
for d in xxz_triangular_19 heisenberg_pyrochlore_32 heisenberg_chain_24 heisenberg_triangular_19; do tools/run_n.sh 1 $d f1_$d --steps 50 --cpu-seconds 5; done
