#!/bin/bash
# one GPU: full GPU suite, ncu full capture of the streaming kernel on heisenberg_chain_40 (ONE GPU), smoke
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/g6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g6_pytest.log
tail -4 gpurun_out/g6_pytest.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cached_matvec_kernel -s 2 -c 1 -o gpurun_out/r02b_prof_cached_chain40_n1 python tools/cached_probe.py heisenberg_chain_40 3 > gpurun_out/g6_ncu_chain40.log 2>&1
tail -2 gpurun_out/g6_ncu_chain40.log | cut -c1-400
ls -la gpurun_out/*chain40*.ncu-rep
