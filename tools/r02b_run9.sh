#!/bin/bash
# one GPU: exit codes and wall time of the driver's default commands; initcheck teardown with / without background NVRTC
mkdir -p gpurun_out
T0=$SECONDS; python bench.py > gpurun_out/g9_bench.json 2> gpurun_out/g9_bench.err; echo "bench.py rc=$? wall $((SECONDS-T0)) s" | tee gpurun_out/g9_rc.txt
T0=$SECONDS; python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/g9_smoke.log 2>&1; echo "smoke rc=$? wall $((SECONDS-T0)) s" | tee -a gpurun_out/g9_rc.txt
tail -1 gpurun_out/g9_smoke.log
export SPED_FILL_CHUNK_BYTES=300
for pf in 0 x; do
  [ $pf = 0 ] && export SPED_JIT_PREFETCH=0 || unset SPED_JIT_PREFETCH
  timeout 300 compute-sanitizer --tool initcheck --error-exitcode 9 python tools/sanitize_target.py heisenberg_chain_10 heisenberg_square_4x4 chain_8_k1_complex > gpurun_out/g9_initcheck_prefetch_$pf.log 2>&1
  echo "initcheck (SPED_JIT_PREFETCH=$pf): rc=$? $(grep -c SANITIZE_TARGET_OK gpurun_out/g9_initcheck_prefetch_$pf.log) decks ok; $(grep -E 'ERROR SUMMARY|terminate' gpurun_out/g9_initcheck_prefetch_$pf.log | tr '\n' ' ')" | tee -a gpurun_out/g9_rc.txt
done
python - <<PY
import json
d=json.load(open("gpurun_out/g9_bench.json")); x=d["extra"]; c=x.get("chain_40",{})
print("6x6", d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], x["time_to_ground_state_cold_s"], x["time_to_ground_state_s"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
print("c40", {k:c.get(k) for k in ["ms_per_step","time_to_ground_state_s","E0","sample_parity_rel_l2","error"]})
PY
