#!/bin/bash
# Sweeps the tuning knobs of the streaming (operator-cache) matvec kernel on one GPU and prints
# kernel ms per setting (f64 and f32 storage).  Run under gpurun; results land in gpurun_out/.
#   tools/sweep_cached.sh DECK "variant:bondorder:l2fetch" ...
mkdir -p gpurun_out
DECK=${1:-heisenberg_square_6x6}
shift
for spec in "$@"; do
  IFS=: read v bo fetch <<< "$spec"
  tag=${DECK}_v${v}_b${bo}_f${fetch:-def}
  SPED_LOG=1 SPED_CACHED_VARIANT=$v SPED_BOND_ORDER=$bo SPED_L2_FETCH=$fetch timeout ${RUN_TIMEOUT:-600} python bench.py --config $DECK --steps 30 --warmup 3 \
    --no-eigh --no-cpu --e2e-host-gb 0 > gpurun_out/sweep_$tag.json 2> gpurun_out/sweep_$tag.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/sweep_$tag.json"))
    f = d["extra"].get("float32_storage") or {}
    print("$tag: f64 %.4f ms frac %.3f | f32 %.4f ms frac %.3f | mf %.1f ms | cache %.2f GB build %.3f s" % (
        d["ms_per_step"], d["roofline"]["frac"], f.get("kernel_ms", 0), f.get("roofline_frac_hbm", 0),
        d["extra"]["matrix_free"]["ms_per_step"], d["extra"]["operator_cache"]["bytes"] / 1e9,
        d["extra"]["operator_cache"]["build_seconds"]))
except Exception as e:
    print("$tag failed:", e)
PY
  grep -h "L2 fetch" gpurun_out/sweep_$tag.err | head -1
done
