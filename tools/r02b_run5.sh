#!/bin/bash
# N GPUs: the driver's bench line with the default exchange (copy engines for long shards), chain_40 traces
N=${1:-8}
mkdir -p gpurun_out
export RUN_TIMEOUT=900
SPED_LOG=1 SPED_OVERLAP_TRACE=12 tools/run_n.sh $N heisenberg_square_6x6 g5_bench_n$N --steps 30 --no-cpu
grep -E "overlapped matvec" gpurun_out/g5_bench_n$N.err | grep -E "rank 0" | tail -6
grep -E "cold eigh|warm eigh|parity|copy-engine" gpurun_out/g5_bench_n$N.err | grep -E "rank 0|x$N|copy" | head
python - <<PY
import json
d = json.load(open("gpurun_out/g5_bench_n$N.json"))
c = d["extra"].get("chain_40", {})
print("6x6 ms/step", d["ms_per_step"], "chain_40 ms/step", c.get("ms_per_step"), "kernel", c.get("kernel_ms"), "ttgs", c.get("time_to_ground_state_s"), "cold", c.get("time_to_ground_state_cold_s"), "E0", c.get("E0"), "parity", c.get("sample_parity_rel_l2"), "build", c.get("basis_build_s"))
print(c.get("eigh_stats"))
PY
