#!/bin/bash
# 8 GPUs: overlap trace of chain_40 with the copy-engine exchange, chain_42 re-measured
mkdir -p gpurun_out
export RUN_TIMEOUT=700
SPED_OVERLAP_TRACE=14 tools/run_n.sh 8 heisenberg_chain_40 g7_c40_trace --steps 10 --no-eigh --no-parity --no-cpu --sharded-deck ''
grep "rank 0 overlapped" gpurun_out/g7_c40_trace.err | tail -4
T0=$SECONDS; SPED_LOG=1 tools/run_n.sh 8 heisenberg_chain_42 g7_c42 --steps 10 --no-cpu --sharded-deck ''; echo "chain_42 wall $((SECONDS-T0)) s"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/g7_c42.json")); x=d["extra"]
    print("chain_42", d["ms_per_step"], x["kernel_ms"], d["roofline"], "build", x["basis_build_s"], "parity", x.get("sample_parity_rel_l2"), x.get("basis_check"), "ttgs", x.get("time_to_ground_state_s"), x.get("time_to_ground_state_cold_s"), x.get("eigenvalues"), x.get("eigh_stats"), x["operator_cache"], x["matrix_free"]["ms_per_step"])
except Exception as e: print("failed", e)
PY
grep -E "rank 0|copy-engine|operator cache" gpurun_out/g7_c42.err | head -12 | cut -c1-250
tail -3 gpurun_out/g7_c42.err | cut -c1-300
