#!/bin/bash
# 8 GPUs: forced-rounds multi-rank test, the driver's bench line (6x6 + chain_40), an overlap trace of chain_40, chain_42
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | sed -n 2p
timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "8-True" > gpurun_out/m8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m8_pytest.log
tail -4 gpurun_out/m8_pytest.log
export RUN_TIMEOUT=700
T0=$SECONDS; tools/run_n.sh 8 heisenberg_square_6x6 m8_bench --steps 50; echo "bench wall $((SECONDS-T0)) s"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/m8_bench.json")); x=d["extra"]
    print("6x6 e2e", d["e2e"], "build", x["basis_build_s"], "parity", x.get("sample_parity_rel_l2"), "ttgs", x.get("time_to_ground_state_s"), x.get("time_to_ground_state_cold_s"), x.get("eigh_stats"))
    print("chain_40", json.dumps(x.get("chain_40"))[:1800])
except Exception as e: print("failed", e)
PY
SPED_OVERLAP_TRACE=14 tools/run_n.sh 8 heisenberg_chain_40 m8_c40_trace --steps 10 --no-eigh --no-parity --sharded-deck ''
grep "rank 0 overlapped" gpurun_out/m8_c40_trace.err | tail -4
T0=$SECONDS; tools/run_n.sh 8 heisenberg_chain_42 m8_c42 --steps 10 --sharded-deck ''; echo "chain_42 wall $((SECONDS-T0)) s"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/m8_c42.json")); x=d["extra"]
    print("chain_42", d["ms_per_step"], x["kernel_ms"], d["roofline"], "build", x["basis_build_s"], "parity", x.get("sample_parity_rel_l2"), x.get("basis_check"), "ttgs", x.get("time_to_ground_state_s"), x.get("time_to_ground_state_cold_s"), x.get("eigenvalues"), x.get("eigh_stats"), x["operator_cache"], x["matrix_free"]["ms_per_step"])
except Exception as e: print("failed", e)
PY
tail -3 gpurun_out/m8_c42.err | cut -c1-300
