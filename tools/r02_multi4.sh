#!/bin/bash
# 4 GPUs: multi-rank tests (two exchange rounds, sharded host-pointer entry), then the driver's bench line
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 700 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/m4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m4_pytest.log
tail -12 gpurun_out/m4_pytest.log
grep -q "rc=0" gpurun_out/m4_pytest.log || exit 1
export RUN_TIMEOUT=600
SPED_OVERLAP_TRACE=6 tools/run_n.sh 4 heisenberg_square_6x6 m4_bench --steps 30
grep "overlapped matvec" gpurun_out/m4_bench.err | head -12
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/m4_bench.json")); x=d["extra"]
    print("6x6 e2e", d["e2e"], "build", x["basis_build_s"], "parity", x.get("sample_parity_rel_l2"), x.get("basis_check"))
    print("chain_40", json.dumps(x.get("chain_40"))[:1500])
except Exception as e: print("failed", e)
PY
