#!/bin/bash
# compute-sanitizer over every kernel family on small decks (race / memory checks).
# Logs land in gpurun_out/sanitize_*.log; a summary line per run is printed.
mkdir -p gpurun_out
DECKS="heisenberg_chain_10 heisenberg_square_4x4 chain_8_k1_complex"
export SPED_FILL_CHUNK_BYTES=${SPED_FILL_CHUNK_BYTES:-300}  # several fill chunks even on these small decks
for tool in memcheck racecheck initcheck synccheck; do
  log=gpurun_out/sanitize_${tool}_1gpu.log
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_target.py $DECKS > $log 2>&1
  echo "$tool 1 GPU: rc=$? $(grep -c SANITIZE_TARGET_OK $log) decks ok; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
done
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  for tool in memcheck racecheck; do
    for xchg in nccl ce; do  # both transports of the shard exchange
      [ "$tool" = racecheck ] && [ "$xchg" = ce ] && continue
      log=gpurun_out/sanitize_${tool}_2gpu_${xchg}.log
      SPED_EXCHANGE=$xchg timeout 900 compute-sanitizer --tool $tool --target-processes all --error-exitcode 9 \
        python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tools/sanitize_target.py $DECKS > $log 2>&1
      echo "$tool 2 GPUs ($xchg exchange): rc=$? $(grep -c SANITIZE_TARGET_OK $log) rank-decks ok; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -2 | tr '\n' ' ')"
    done
  done
fi
