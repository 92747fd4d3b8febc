#!/bin/bash
# N GPUs (argument): copy-engine exchange (SPED_EXCHANGE=ce) against NCCL -- tests, then the bench line of both
N=${1:-2}
mkdir -p gpurun_out
SPED_EXCHANGE=ce timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/g4_pytest_ce_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g4_pytest_ce_n$N.log
tail -5 gpurun_out/g4_pytest_ce_n$N.log
export RUN_TIMEOUT=900
SPED_EXCHANGE=ce SPED_LOG=1 SPED_OVERLAP_TRACE=6 tools/run_n.sh $N heisenberg_square_6x6 g4_bench_ce_n$N --steps 30 --no-cpu
grep -E "overlapped matvec|copy-engine|cold eigh|warm eigh|parity" gpurun_out/g4_bench_ce_n$N.err | grep -E "rank 0|copy-engine|x$N" | head -24
python - <<PY
import json
d = json.load(open("gpurun_out/g4_bench_ce_n$N.json"))
c = d["extra"].get("chain_40", {})
print("CE  6x6 ms/step", d["ms_per_step"], "chain_40 ms/step", c.get("ms_per_step"), "kernel", c.get("kernel_ms"), "ttgs", c.get("time_to_ground_state_s"), "E0", c.get("E0"), "parity", c.get("sample_parity_rel_l2"))
PY
if [ "$2" = "both" ]; then
SPED_EXCHANGE=nccl SPED_OVERLAP_TRACE=6 tools/run_n.sh $N heisenberg_square_6x6 g4_bench_nccl_n$N --steps 30 --no-cpu
grep -E "overlapped matvec" gpurun_out/g4_bench_nccl_n$N.err | grep -E "rank 0" | head -12
python - <<PY
import json
d = json.load(open("gpurun_out/g4_bench_nccl_n$N.json"))
c = d["extra"].get("chain_40", {})
print("NCCL 6x6 ms/step", d["ms_per_step"], "chain_40 ms/step", c.get("ms_per_step"), "kernel", c.get("kernel_ms"), "ttgs", c.get("time_to_ground_state_s"), "E0", c.get("E0"), "parity", c.get("sample_parity_rel_l2"))
PY
fi
