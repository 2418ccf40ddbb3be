#!/bin/bash
# compute-sanitizer passes over the path-1 kernels at small shapes (memcheck: out-of-bounds / misaligned accesses in global
# and shared memory; racecheck: shared-memory hazards between the lanes / warps of a CTA; synccheck: barrier misuse).
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_sanitizer.sh'
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SEL='test_kernel_vs_fp64_oracle or test_optional_inputs or test_mode_backward or test_fused_colourisation_matches'
for tool in memcheck racecheck synccheck; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 3 \
        python -m pytest tests/test_gpu_template.py -q -x -m gpu -k "$SEL" > gpurun_out/sanitizer_$tool.txt 2>&1
    echo "$tool rc=$? $(grep -E 'passed|failed|error' gpurun_out/sanitizer_$tool.txt | tail -1) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$tool.txt | tail -1)"
done
