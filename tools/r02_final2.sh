#!/bin/bash
# one GPU: full GPU suite again (conformance fix, parallel bucket table), launch list covering the timed region,
# sector cross-check of the 6x6 energy, ncu capture of the chain_36 kernel (gathers scattered over 504 MB)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/f2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f2_pytest.log
tail -4 gpurun_out/f2_pytest.log
timeout 300 python tools/eigh_probe.py heisenberg_square_6x6 2>&1 | grep -E "first|second|dropped:"
timeout 600 python tools/sector_cross_check.py heisenberg_square_6x6 > gpurun_out/r02_sector_cross_check_6x6.json 2> gpurun_out/f2_cross.err; cat gpurun_out/r02_sector_cross_check_6x6.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_6x6.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-parity --no-eigh > gpurun_out/f2_ncu_list.log 2>&1
tail -1 gpurun_out/f2_ncu_list.log | cut -c1-160
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cached_matvec_kernel -s 4 -c 1 -o gpurun_out/r02_prof_cached_chain36 python bench.py --config heisenberg_chain_36 --steps 3 --warmup 3 --no-eigh --no-cpu --no-parity --e2e-host-gb 0 > gpurun_out/f2_ncu_c36.log 2>&1
export RUN_TIMEOUT=600
tools/run_n.sh 1 heisenberg_square_6x6 f2_bench --steps 50
tools/run_n.sh 1 heisenberg_chain_36 f2_chain36 --steps 30 --no-cpu
