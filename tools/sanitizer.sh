#!/bin/bash
# compute-sanitizer over the kernels with mbarrier pipelines, TMEM, clusters / distributed shared memory and cp.async
# staging: memcheck + racecheck + synccheck on small shapes.  Summaries -> gpurun_out/ (copied to profiles/).
mkdir -p gpurun_out
TESTS="tests/test_gpu_tc.py::test_tc_splitk_cluster_reduction tests/test_gpu_tc.py::test_halo_strip_conv_matches_fp32 tests/test_gpu_tc.py::test_groupnorm_cluster_fp16_storage tests/test_gpu_tc_bwd.py::test_tc_wgrad_and_dgrad_match_torch tests/test_gpu_tc_bwd.py::test_groupnorm_cluster_backward_matches_torch tests/test_gpu_tc.py::test_tc_gemm_3xtf32_is_fp32_accurate tests/test_gpu_audio.py"
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 python -m pytest $TESTS -m gpu -x -q > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "== $tool: $(grep -c 'passed' gpurun_out/r02_sanitizer_$tool.log) summary lines"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error:|Race reported|Hazard" gpurun_out/r02_sanitizer_$tool.log | sort | uniq -c | sort -rn | head -12
done
