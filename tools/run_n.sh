#!/bin/bash
# tools/run_n.sh N DECK TAG [bench args...]: bench.py on N GPUs of this box, JSON line to gpurun_out/TAG.json.
# Every run is wrapped in `timeout` (RUN_TIMEOUT seconds, default 600): a mismatched collective hangs
# all ranks, and an unbounded hang burns N x the GPU budget.
N=$1; DECK=$2; TAG=$3; shift 3
mkdir -p gpurun_out
if [ "$N" = 1 ]; then
  timeout ${RUN_TIMEOUT:-600} python bench.py --config $DECK "$@" > gpurun_out/$TAG.json 2> gpurun_out/$TAG.err
else
  timeout ${RUN_TIMEOUT:-600} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --config $DECK "$@" > gpurun_out/$TAG.json 2> gpurun_out/$TAG.err
fi
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/$TAG.json"))
    x = d["extra"]
    print("$TAG: %.4f ms/step kernel %s frac %.3f value %.3e e2e %s ttgs %s E0 %s mv %s" % (
        d["ms_per_step"], x.get("kernel_ms_per_rank"), d["roofline"]["frac"], d["value"], (d.get("e2e") or {}).get("value"),
        x.get("time_to_ground_state_s"), (x.get("eigenvalues") or [None])[0], x.get("eigh_matvecs")))
except Exception as e:
    print("$TAG failed:", e)
    import subprocess
    print(subprocess.run(["tail", "-5", "gpurun_out/$TAG.err"], capture_output=True, text=True).stdout)
PY
