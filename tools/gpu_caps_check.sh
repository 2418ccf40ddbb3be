#!/bin/bash
# One gpurun call for the path-2 kernels: parity tests, kernel micro-benchmarks of the variants, one ncu capture.
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
log() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a gpurun_out/timeline.txt; }
: > gpurun_out/timeline.txt
log "start $(nvidia-smi -L | head -1)"
timeout 400 python -m pytest tests/test_gpu_capsule.py -q --tb=short -m gpu -x > gpurun_out/pytest_caps.txt 2>&1
log "pytest_caps rc=$? $(tail -1 gpurun_out/pytest_caps.txt)"
: > gpurun_out/kernel_bench.jsonl
run() {  # label, env...
    local label=$1; shift
    echo "# $label" >> gpurun_out/kernel_bench.jsonl
    env "$@" timeout 300 python tools/kernel_bench.py --configs "${CFGS:-mnist32}" --batches "${BATCHES:-1024,8192}" --iters 10 --only caps \
        >> gpurun_out/kernel_bench.jsonl 2>> gpurun_out/kernel_bench.err
    log "kernel_bench $label rc=$?"
}
run v3_default X=1
run v3_np4x2 SCAE_CAPS3_NP=4 SCAE_CAPS3_MINB=2
run v3_np2_s2 SCAE_CAPS3_NP=2 SCAE_CAPS3_STAGES=2
run v2 SCAE_CAPS_IMPL=v2
CFGS=mnist10,stress,color BATCHES=8192 run v3_other X=1
CFGS=mnist10,stress,color BATCHES=8192 run v2_other SCAE_CAPS_IMPL=v2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:caps3_fwd -s 2 -c 2 -o gpurun_out/caps3_fwd \
    python tools/kernel_bench.py --configs mnist32 --batches 8192 --iters 2 --only caps > gpurun_out/ncu_caps3.log 2>&1
log "ncu rc=$?"
log done
