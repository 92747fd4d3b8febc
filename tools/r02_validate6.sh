#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/v6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v6_pytest.log
tail -8 gpurun_out/v6_pytest.log
for d in heisenberg_square_6x6 xxz_triangular_19 heisenberg_chain_24 heisenberg_pyrochlore_32 heisenberg_triangular_19; do
  timeout 300 python tools/eigh_probe.py $d 2>&1 | grep -v "^$" | grep -E "first|second|dropped:" 
done
