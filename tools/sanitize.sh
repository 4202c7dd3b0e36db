#!/usr/bin/env bash
# compute-sanitizer passes over the kernel tests (run on a GPU box; slow: use the small-shape tests only).
# NOTE: this pool's gpurun refuses compute-sanitizer ("closed on this pool", rc 86, gpurun_out/r2g_memcheck_dense.log),
# so there is no sanitizer log under profiles/; bounds are covered by the out-of-range / ragged / empty-input tests.
#   gpurun --timeout 900 -- 'bash tools/sanitize.sh'
# Writes gpurun_out/sanitize_{memcheck,racecheck,initcheck}.log; exit code != 0 if any tool reports an error.
set -u
rc=0
TESTS="tests/test_gpu_spmm.py tests/test_gpu_modules.py tests/test_gpu_graph.py"
for tool in memcheck racecheck initcheck; do
  compute-sanitizer --tool "$tool" --error-exitcode 9 --launch-timeout 0 \
    python -m pytest $TESTS -x -q -m gpu -k "not bench and not property" \
    > "gpurun_out/sanitize_${tool}.log" 2>&1 || rc=$?
  tail -3 "gpurun_out/sanitize_${tool}.log"
  grep -c "ERROR SUMMARY: 0 errors" "gpurun_out/sanitize_${tool}.log" || true
done
exit $rc
