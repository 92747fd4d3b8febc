#!/bin/bash
# two GPUs: multi-GPU tests, the driver's 2-GPU bench line (6x6 + chain_40), conformance driver exit-race loop
mkdir -p gpurun_out
gcc -std=c11 -I include tests/conformance.c -o /tmp/conformance -L spin-ed_b200/lib -lsped -lm -Wl,-rpath,$PWD/spin-ed_b200/lib
ok=0; for i in 1 2 3 4 5 6 7 8; do SPED_CACHE_DIR= CUDA_VISIBLE_DEVICES=0 /tmp/conformance > /tmp/conf.out 2>&1; rc=$?; [ $rc = 0 ] && ok=$((ok+1)) || echo "conformance run $i rc=$rc"; done; echo "conformance (cold cubin cache): $ok of 8 runs ok" | tee gpurun_out/g3_conformance.txt
ok=0; for i in 1 2 3 4; do CUDA_VISIBLE_DEVICES=0 /tmp/conformance > /tmp/conf.out 2>&1; rc=$?; [ $rc = 0 ] && ok=$((ok+1)) || echo "conformance run $i rc=$rc"; done; echo "conformance (warm cubin cache): $ok of 4 runs ok" | tee -a gpurun_out/g3_conformance.txt
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_conformance.py -m gpu -q > gpurun_out/g3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g3_pytest.log
tail -6 gpurun_out/g3_pytest.log
export RUN_TIMEOUT=900
SPED_LOG=1 tools/run_n.sh 2 heisenberg_square_6x6 g3_bench_n2 --steps 30
grep -E "cold eigh|warm eigh|operator cache|parity|jit" gpurun_out/g3_bench_n2.err | head -30
