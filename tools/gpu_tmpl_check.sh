#!/bin/bash
# Path-1 backward iteration in one gpurun call: parity of the template path, micro-benchmarks (L2 flushed), one
# --set full capture of the backward kernel, then the whole-step bench.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_tmpl_check.sh'
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
log() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a gpurun_out/timeline.txt; }
: > gpurun_out/timeline.txt
log "start $(nvidia-smi -L | head -1)"
timeout 400 python -m pytest tests/test_gpu_template.py -q --tb=short -m gpu -x > gpurun_out/pytest_tmpl.txt 2>&1
log "pytest_tmpl rc=$? $(tail -1 gpurun_out/pytest_tmpl.txt)"
timeout 300 python tools/kernel_bench.py --configs mnist32,stress,color --batches 1024,8192 --iters 10 --only tmpl \
    > gpurun_out/kernel_bench_tmpl.jsonl 2> gpurun_out/kernel_bench_tmpl.err
log "kernel_bench rc=$?"
for v in $TMPL_VARIANTS; do
    env $v timeout 300 python tools/kernel_bench.py --configs mnist32,stress --batches 1024 --iters 10 --only tmpl \
        > gpurun_out/kernel_bench_tmpl_$v.jsonl 2> gpurun_out/kernel_bench_tmpl_$v.err
    log "kernel_bench $v rc=$?"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'tmpl_ll_bwd' -s 3 -c 1 -o gpurun_out/prof_tmpl_bwd \
    python tools/kernel_bench.py --configs mnist32 --batches 1024 --iters 2 --only tmpl > gpurun_out/ncu_tmpl.log 2>&1
log "ncu tmpl rc=$?"
timeout 300 python -m pytest tests/test_gpu_model.py -q --tb=short -m gpu -x > gpurun_out/pytest_model.txt 2>&1
log "pytest_model rc=$? $(tail -1 gpurun_out/pytest_model.txt)"
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
log "bench rc=$?"
log done
