mkdir -p gpurun_out
timeout 130 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden_teacher_forced or golden_inference_ragged or (fused_gemm_layernorm and 3xtf32) or (attention_vs_torch and 3xtf32) or tap_gemm_3xtf32" > gpurun_out/memcheck_final2.log 2>&1
grep -E "passed|failed|ERROR SUMMARY" gpurun_out/memcheck_final2.log | tail -3
