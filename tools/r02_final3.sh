#!/bin/bash
# 2 GPUs: conformance tests again, 2-GPU sanitizer runs, multi-rank tests at 2 (applyLocal included)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conformance.py tests/test_gpu_driver.py -m gpu -q > gpurun_out/f3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f3_pytest.log
tail -3 gpurun_out/f3_pytest.log
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "2-" > gpurun_out/f3_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f3_pytest_multi.log
tail -3 gpurun_out/f3_pytest_multi.log
DECKS="heisenberg_chain_10 heisenberg_square_4x4 chain_8_k1_complex"
for tool in memcheck racecheck; do
  log=gpurun_out/sanitize_${tool}_2gpu.log
  timeout 600 compute-sanitizer --tool $tool --target-processes all --error-exitcode 9 \
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tools/sanitize_target.py $DECKS > $log 2>&1
  echo "$tool 2 GPUs: rc=$? $(grep -c SANITIZE_TARGET_OK $log) rank-decks ok; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | sort | uniq -c | tr '\n' ' ')"
done
