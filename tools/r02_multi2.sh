#!/bin/bash
# 2 GPUs: multi-rank tests, then the driver's bench line (6x6 + chain_40)
mkdir -p gpurun_out
export RUN_TIMEOUT=800
T0=$SECONDS; tools/run_n.sh 2 heisenberg_square_6x6 m2_bench --steps 50; echo "bench wall $((SECONDS-T0)) s"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/m2_bench.json")); x=d["extra"]
    print("6x6 e2e", d["e2e"], "build", x["basis_build_s"], "parity", x.get("sample_parity_rel_l2"), "ttgs", x.get("time_to_ground_state_s"), x.get("time_to_ground_state_cold_s"), x.get("eigh_stats"))
    print("chain_40", json.dumps(x.get("chain_40"))[:1800])
except Exception as e: print("failed", e)
PY
tail -5 gpurun_out/m2_bench.err
