#!/bin/bash
# two GPUs: full GPU suite (multi-GPU cases at world 2 included), initcheck, the driver's 2-GPU bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/g8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g8_pytest.log
tail -6 gpurun_out/g8_pytest.log
export SPED_FILL_CHUNK_BYTES=300
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 9 python tools/sanitize_target.py heisenberg_chain_10 heisenberg_square_4x4 chain_8_k1_complex > gpurun_out/sanitize_initcheck_1gpu.log 2>&1
echo "initcheck 1 GPU: rc=$? $(grep -c SANITIZE_TARGET_OK gpurun_out/sanitize_initcheck_1gpu.log) decks ok; $(grep -E 'ERROR SUMMARY' gpurun_out/sanitize_initcheck_1gpu.log | tail -1)"
timeout 600 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tools/sanitize_target.py heisenberg_square_4x4 chain_8_k1_complex > gpurun_out/sanitize_memcheck_2gpu_staged.log 2>&1
echo "memcheck 2 GPUs (staged fill): rc=$? $(grep -c SANITIZE_TARGET_OK gpurun_out/sanitize_memcheck_2gpu_staged.log) rank-decks ok; $(grep -E 'ERROR SUMMARY' gpurun_out/sanitize_memcheck_2gpu_staged.log | tail -2 | tr '\n' ' ')"
unset SPED_FILL_CHUNK_BYTES
export RUN_TIMEOUT=900
SPED_LOG=1 tools/run_n.sh 2 heisenberg_square_6x6 g8_bench_n2 --steps 30 --no-cpu
grep -E "cold eigh|warm eigh|operator cache|parity" gpurun_out/g8_bench_n2.err | grep -E "rank 0|x2|operator cache" | head -14
