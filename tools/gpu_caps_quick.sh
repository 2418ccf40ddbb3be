#!/bin/bash
# Quick path-2 iteration on the GPU: parity of the fast path, micro-benchmark of the listed variants, one ncu capture.
#   VARIANTS="label:ENV=1,ENV2=2 label2:" bash tools/gpu_caps_quick.sh
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
log() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a gpurun_out/timeline.txt; }
: > gpurun_out/timeline.txt
timeout 400 python -m pytest tests/test_gpu_capsule.py -q --tb=short -m gpu -x > gpurun_out/pytest_caps.txt 2>&1
log "pytest_caps rc=$? $(tail -1 gpurun_out/pytest_caps.txt)"
: > gpurun_out/kernel_bench.jsonl
for spec in ${VARIANTS:-default:X=1}; do
    label=${spec%%:*}; envs=${spec#*:}
    echo "# $label" >> gpurun_out/kernel_bench.jsonl
    env ${envs//,/ } timeout 300 python tools/kernel_bench.py --configs "${CFGS:-mnist32}" --batches "${BATCHES:-1024,8192}" --iters 10 --only caps \
        >> gpurun_out/kernel_bench.jsonl 2>> gpurun_out/kernel_bench.err
    log "kernel_bench $label rc=$?"
done
if [ -n "$NCU" ]; then
    env ${NCU_ENV//,/ } timeout 300 ncu --set full --clock-control none --import-source on -k regex:$NCU -s 2 -c 1 -o gpurun_out/$NCU \
        python tools/kernel_bench.py --configs mnist32 --batches 8192 --iters 2 --only caps > gpurun_out/ncu_$NCU.log 2>&1
    log "ncu rc=$?"
fi
log done
