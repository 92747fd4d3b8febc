#!/bin/bash
# one GPU: full GPU suite, the driver's bench line (6x6 + chain_40 on ONE GPU), ncu launch list + full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,memory.used --format=csv > gpurun_out/g1_smi.txt 2>&1
free -g >> gpurun_out/g1_smi.txt; nproc >> gpurun_out/g1_smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/g1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g1_pytest.log
tail -15 gpurun_out/g1_pytest.log
export RUN_TIMEOUT=900
SPED_LOG=1 tools/run_n.sh 1 heisenberg_square_6x6 g1_bench --steps 50
tail -40 gpurun_out/g1_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02b_launches_6x6.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-parity --no-sharded-at-one > gpurun_out/g1_ncu_list.log 2>&1
tail -1 gpurun_out/g1_ncu_list.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cached_matvec_kernel -s 4 -c 1 -o gpurun_out/r02b_prof_cached_6x6 python bench.py --steps 3 --warmup 3 --no-eigh --no-cpu --no-parity --e2e-host-gb 0 --no-sharded-at-one > gpurun_out/g1_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
