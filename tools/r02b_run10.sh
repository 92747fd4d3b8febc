#!/bin/bash
mkdir -p gpurun_out
export SPED_FILL_CHUNK_BYTES=300
for v in nojit default; do
  [ $v = nojit ] && export SPED_JIT=0 || unset SPED_JIT
  timeout 300 compute-sanitizer --tool initcheck --error-exitcode 9 python -X faulthandler tools/sanitize_target.py heisenberg_chain_10 chain_8_k1_complex > gpurun_out/g10_initcheck_$v.log 2>&1
  echo "initcheck ($v): rc=$? $(grep -c SANITIZE_TARGET_OK gpurun_out/g10_initcheck_$v.log) decks ok"; grep -v "^=========  *Host Frame" gpurun_out/g10_initcheck_$v.log | tail -25
done
