mkdir -p gpurun_out
timeout 170 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "(golden_teacher_forced and 3xtf32) or (fused_gemm_layernorm and 3xtf32-384 and 51-1024)" > gpurun_out/memcheck_final.log 2>&1
grep -E "passed|failed|ERROR SUMMARY" gpurun_out/memcheck_final.log | tail -3
timeout 80 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden_teacher_forced and 3xtf32" > gpurun_out/synccheck_final.log 2>&1
grep -E "passed|failed|ERROR SUMMARY" gpurun_out/synccheck_final.log | tail -3
