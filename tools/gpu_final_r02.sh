#!/bin/bash
# Round-2 evidence in one gpurun call: the whole GPU test suite, smoke, bench.py, kernel micro-benchmarks at every BASELINE
# shape (L2 flushed), the ncu launch list of one step and one --set full capture of the four likelihood kernels.
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
log() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a gpurun_out/timeline.txt; }
: > gpurun_out/timeline.txt
log "start $(nvidia-smi -L | head -1)"
timeout 900 python -m pytest tests -q --tb=short -m gpu > gpurun_out/pytest_gpu.txt 2>&1
log "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu.txt)"
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.txt 2>&1
log "smoke rc=$? $(tail -1 gpurun_out/smoke.txt)"
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
log "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
log "bench reference rc=$?"
timeout 400 python tools/kernel_bench.py --configs mnist32,mnist10,stress,color --batches 1024,8192 --iters 10 > gpurun_out/kernel_bench.jsonl 2> gpurun_out/kernel_bench.err
log "kernel_bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
log "ncu launches rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'tmpl_ll|caps' \
    -o gpurun_out/prof_step python tools/profile_step.py > gpurun_out/ncu_step.log 2>&1
log "ncu step rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'caps3' -s 4 -c 2 -o gpurun_out/prof_caps8192 \
    python tools/kernel_bench.py --configs mnist32 --batches 8192 --iters 2 --only caps > gpurun_out/ncu_caps.log 2>&1
log "ncu caps rc=$?"
log done
