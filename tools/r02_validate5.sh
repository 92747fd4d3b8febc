#!/bin/bash
# block eigensolver: GPU tests, then solver timings on the decks
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/v5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v5_pytest.log
tail -15 gpurun_out/v5_pytest.log
export RUN_TIMEOUT=600
tools/run_n.sh 1 heisenberg_square_6x6 v5_6x6 --steps 30 --no-cpu --no-parity
tools/run_n.sh 1 xxz_triangular_19 v5_xxz --steps 30 --no-cpu --no-parity
tools/run_n.sh 1 heisenberg_pyrochlore_32 v5_pyro --steps 30 --no-cpu --no-parity
tools/run_n.sh 1 heisenberg_chain_24 v5_chain24 --steps 30 --no-cpu --no-parity
tools/run_n.sh 1 heisenberg_triangular_19 v5_tri19 --steps 30 --no-cpu --no-parity
for t in v5_6x6 v5_xxz v5_pyro v5_chain24 v5_tri19; do python - <<PY
import json
try:
    x=json.load(open("gpurun_out/$t.json"))["extra"]; print("$t", x.get("time_to_ground_state_cold_s"), x.get("time_to_ground_state_s"), x.get("eigenvalues"), x.get("eigh_stats"))
except Exception as e: print("$t", e)
PY
done
