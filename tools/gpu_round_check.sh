#!/bin/bash
# One gpurun call, most valuable first (a call that runs out of budget still leaves what it finished in gpurun_out/):
#   1 parity tests of the newest code  2 bench.py  3 A/B of the environment switches  4 ncu launch list of one step
#   5 the remaining GPU tests  6 smoke()
#   /usr/local/graft/bin/gpurun --timeout 560 -- 'bash tools/gpu_round_check.sh [quick]'     (quick: steps 1-3 only; final: steps 1, 2 and 4)
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
log() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a gpurun_out/timeline.txt; }
: > gpurun_out/timeline.txt
log "start $(nvidia-smi -L | head -1)"
timeout 300 python -m pytest tests/test_gpu_plumbing.py tests/test_gpu_graph.py tests/test_gpu_model.py -q --tb=short -m gpu \
    > gpurun_out/pytest_new.txt 2>&1
log "pytest_new rc=$? $(tail -1 gpurun_out/pytest_new.txt)"
timeout 240 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
log "bench rc=$?"
if [ "$1" != final ]; then
    timeout 180 python tools/ab_step.py > gpurun_out/ab_step.jsonl 2> gpurun_out/ab_step.err
    log "ab_step rc=$?"
fi
[ "$1" = quick ] && { log "quick: done"; exit 0; }
timeout 180 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
log "ncu launches rc=$?"
[ "$1" = final ] && { log "final: done"; exit 0; }
timeout 400 python -m pytest tests/test_gpu_template.py tests/test_gpu_capsule.py -q --tb=short -m gpu \
    > gpurun_out/pytest_rest.txt 2>&1
log "pytest_rest rc=$? $(tail -1 gpurun_out/pytest_rest.txt)"
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.txt 2>&1
log "smoke rc=$?"
timeout 120 python tools/kernel_bench.py --configs mnist32 --batches 1024,8192 --iters 10 > gpurun_out/kernel_bench.jsonl 2>&1
log "kernel_bench rc=$?"
log done
