#!/usr/bin/env python
"""SCAE train-step benchmark on synthetic MNIST-shaped data (BASELINE.json metric: train images/sec; likelihood-kernel
achieved GB/s vs the measured HBM peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config mnist32|mnist10|stress|color]   # the fused sm_100a path
    python bench.py --impl reference [--steps K] [--warmup W]     # the UNMODIFIED reference (baseline/_ref) on the host cores

A "step" = forward + SCAE.loss + backward + (gradient all-reduce) + RMSprop update on one batch of 1024 images per GPU
(BASELINE.json configs[1]; weak scaling: 8 GPUs = configs[2]'s global batch 8192).  Timing: CUDA events on the
launching stream bracketed by barrier + synchronize, max over ranks.  Prints ONE JSON line on rank 0.  `extra` carries the
other BASELINE configs (whole-step images/s of the 10-capsule, likelihood-stress and colour models), the class-default
vote_type='soft' step, the strong-scaling point at global batch 8192 (N > 1) and the reference moved to the same GPU
("gpu_before"); `cpu_baseline` the reference on the box's host cores with its two hot-path regions timed in isolation.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'scae_train_images_per_sec'
UNIT = 'images/s'


def model_params(n_obj_caps):
    """MNIST SCAE default = the reference's configs/model/mnist.yaml (reconstruct_alternatives: false)."""
    return dict(image_shape=(1, 40, 40), n_classes=10, n_part_caps=40, n_obj_caps=n_obj_caps,
                scae_params=dict(reconstruct_alternatives=False))


# BASELINE.json configs: name -> (factory parameters, kernel shape, description)
CONFIGS = {
    'mnist32': (model_params(32), dict(M=40, C=1, h=11, w=11, H=40, W=40, O=32),
                'MNIST SCAE default (mnist.yaml: 1x40x40, 40 part caps, 32 obj caps, 11x11 templates)'),
    'mnist10': (model_params(10), dict(M=40, C=1, h=11, w=11, H=40, W=40, O=10),
                "MNIST SCAE with BASELINE.json's parenthetical 10 obj caps (1x40x40, 40 part caps, 11x11 templates)"),
    'stress': (dict(image_shape=(1, 64, 64), n_classes=10, n_part_caps=64, n_obj_caps=32,
                    pcae_template_generator_params=dict(template_size=(21, 21)),
                    scae_params=dict(reconstruct_alternatives=False)),
               dict(M=64, C=1, h=21, w=21, H=64, W=64, O=32),
               'likelihood-stress SCAE (BASELINE configs[3]: 1x64x64, 64 part caps, 32 obj caps, 21x21 templates)'),
    'color': (dict(image_shape=(3, 32, 32), n_classes=10, n_part_caps=24, n_obj_caps=32,
                   scae_params=dict(reconstruct_alternatives=False)),
              dict(M=24, C=3, h=11, w=11, H=32, W=32, O=32),
              'SVHN/CIFAR-shaped colour SCAE (BASELINE configs[4]: 3x32x32, 24 part caps, 32 obj caps, 11x11 templates)'),
}


def config_of(args):
    if args.config:
        return args.config
    return 'mnist10' if args.n_obj_caps == 10 else 'mnist32'


def workload_name(config, batch):
    return f'{CONFIGS[config][2]}, batch {batch} per GPU, fwd+loss+bwd+RMSprop, synthetic data'


# ---------------------------------------------------------------------------------------------------------------------
# algorithmic HBM bytes per image (SURVEY.md section 8d / DESIGN.md section 5)
# ---------------------------------------------------------------------------------------------------------------------
def algorithmic_bytes(M=40, C=1, h=11, w=11, H=40, W=40, O=32, fused_color=False):
    """``fused_color``: the train step passes (raw templates, colours) to path 1: per image
    the kernels then read M*C colours instead of M*C*h*w coloured texels and write M*C colour gradients instead of the
    (M,C,h,w) template gradient (the batch-shared raw templates and their gradient are L2-resident, like alpha)."""
    V, A = M, 8 * M + 7
    tex = M * C if fused_color else M * C * h * w
    tmpl_in = 4 * (tex + 6 * M + M + C * H * W)                           # templates | colours, pose, presence, x
    f1 = tmpl_in + 4 * C * H * W + 4 * 2 * C * H * W + 4                  # + log_prob, lse cache, ll
    b1 = tmpl_in + 4 * C * H * W + 4 * 2 * C * H * W + 4 * (tex + 7 * M)  # + grad_out, cache, g_templates|g_color/pose/presence
    caps_in = 4 * O * A + 4 * O * (V + 1) + 4 * 7 * V                     # all_param, noises, x + presence
    # API-complete forward: every tensor of the reference's result dict
    f2 = caps_in + 4 * (O * V * (6 + 7) + 2 * (O + 1) * V + 2 * O + O + V * (6 + 1 + 6 + 1 + 1) + 2) + 8 * 2 * V
    # backward: inputs + saved posterior/lse + upstream (posterior, caps_presence, ll, reg) + g_all_param, written once
    # (ReLU mask, regulariser and the batch sums for cpr_static / biases are fused into the kernel)
    b2 = caps_in + 4 * (O * V + V) + 4 * (O * V + O + 2) + 4 * O * A
    return dict(scae_tmpl_ll_fwd=f1, scae_tmpl_ll_bwd=b1, scae_caps_ll_fwd=f2, scae_caps_ll_bwd=b2)


# Path 1 is bound by the SM issue rate, not by HBM: work units = (M+1) * H * W (pixel, component)
# pairs per image and an ALGORITHMIC lane-instruction count per unit -- 50 for the forward (coordinates, floor/weights,
# four taps, two bilinear blends, two streaming-logsumexp updates; SURVEY's figure) and 120 for the backward: the SASS
# of tmpl_ll_bwd_run_kernel<1,1> issues 81 instructions per (pixel, template) on the path that every pixel takes (record
# and coordinate loads, clamp / floor, the two difference-form samples with their coordinate derivatives, the two
# responsibilities, pose / presence sums, 8 FFMAs of cell sums) plus the cell-change handling, which is about 30 when
# amortised perfectly over a warp (one queue append per cell visit, one conflict-free drain per 32 of them).  The rest
# of what ncu counts (about 180 per pixel and template in profiles/r02b_*: turn-taking for repeated cells, staging, atlas
# flushes) is overhead this figure leaves out.  t_issue = units * instr / (SMs * 128 lanes * f_clk) is the time at a
# perfect issue rate; `measured_issue_ms` below uses the instruction count ncu measured for this build instead.
ISSUE_LANE_INSTR = dict(scae_tmpl_ll_fwd=50, scae_tmpl_ll_bwd=120)


def issue_roof_ms(name, batch, sm_mhz, M=40, H=40, W=40, C=1, sms=148, **_):
    units = (M + 1) * H * W * C * batch
    return units * ISSUE_LANE_INSTR[name] / (sms * 128 * sm_mhz * 1e6) * 1e3


def measured_issue_ms(name, batch, sm_mhz, sms=148):
    """Time the kernel's ACTUAL instruction stream would take at a perfect issue rate: lane-instructions per launch as
    ncu counted them for this build (profiles/issue_counts.json, written by tools/summarize_ncu.py from the committed
    --set full capture; MNIST config), scaled to the batch.  t_measured / this = the issue-pipe utilisation."""
    path = os.path.join(ROOT, 'profiles', 'issue_counts.json')
    if not os.path.exists(path):
        return None
    with open(path) as f:
        rec = json.load(f)
    k = rec.get('kernels', {}).get(name)
    if not k:
        return None
    lane_instr = k['lane_instr_per_launch'] * batch / rec['batch']
    return lane_instr / (sms * 128 * sm_mhz * 1e6) * 1e3


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture, if any."""
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {}


# ---------------------------------------------------------------------------------------------------------------------
# clocks sampling
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons of one GPU, sampled while the timed regions run: through NVML in-process (a few
    samples per 100 ms) when pynvml can find the device, else through `nvidia-smi` (about one sample per second)."""
    FIELDS = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    NAMES = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False      # samples: (sm_mhz, max_mhz, [reason flags])
        self.source, self.nvml, self.handle = 'nvidia-smi', None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = 'GPU-' + str(torch.cuda.get_device_properties(index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml, self.source = pynvml, 'nvml'
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n, h = self.nvml, self.handle
        sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(h)
        masks = (n.nvmlClocksEventReasonHwSlowdown, n.nvmlClocksEventReasonHwThermalSlowdown,
                 n.nvmlClocksEventReasonSwThermalSlowdown, n.nvmlClocksEventReasonSwPowerCap)
        return int(sm), int(mx), [bool(r & m) for m in masks]

    def _sample_smi(self):
        out = subprocess.run(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.FIELDS}',
                              '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
        parts = [p.strip() for p in out.strip().split(',')]
        if len(parts) < 6 or not parts[0].isdigit():
            return None
        return int(parts[0]), int(parts[1]) if parts[1].isdigit() else None, \
            [p.lower().startswith('active') for p in parts[2:6]]

    def run(self):
        while not self.stop_flag:
            try:
                s = self._sample_nvml() if self.nvml is not None else self._sample_smi()
                if s is not None:
                    self.samples.append(s)
            except Exception:
                if self.nvml is not None:                  # NVML query failed: fall back to nvidia-smi
                    self.nvml, self.source = None, 'nvidia-smi'
            time.sleep(0.05 if self.nvml is not None else 0.2)

    def summary(self):
        self.stop_flag = True
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0, source=self.source)
        sm = sorted(s[0] for s in self.samples)
        reasons = [n for i, n in enumerate(self.NAMES) if any(s[2][i] for s in self.samples)]
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=self.samples[0][1], reasons=reasons, samples=len(self.samples),
                    source=self.source)


# ---------------------------------------------------------------------------------------------------------------------
# Reference arm: the UNMODIFIED reference (pip-installed into baseline/_ref by baseline/install_ref.sh, imported through
# baseline/ref_loader.py) driven through its own public API -- factory.make_scae, SCAE.forward, SCAE.loss, RMSprop with
# the experiment's hyper-parameters (base_experiment.py:47-53) -- on the host cores or, for the GPU "before" number, on
# the same B200.  When baseline/_ref is absent (a checkout nobody ran build() on) the oracle port stands in and says so.
# ---------------------------------------------------------------------------------------------------------------------
def reference_stepper(config, batch, device):
    """-> (step function, kind, model-or-None): one train step of the reference (kind 'reference') or the port ('port')."""
    params, shape, _ = CONFIGS[config]
    C, H, W = params['image_shape']
    torch.manual_seed(42)
    image = torch.rand(batch, C, H, W, device=device)
    label = torch.randint(0, 10, (batch,), device=device)
    try:
        from baseline import ref_loader
        ref_loader.load_reference()
        from torch_scae import factory as ref_factory       # the reference's own factory (baseline/_ref)
        model = ref_factory.make_scae(params).to(device).train()
        opt = torch.optim.RMSprop(model.parameters(), lr=3e-5, momentum=0.9, eps=1e-2 / float(batch) ** 2)

        def step():
            opt.zero_grad(set_to_none=True)
            res = model(image=image)
            loss, _ = model.loss(res, image, label)
            loss.backward()
            opt.step()
        return step, 'reference', (model, image, label)
    except ImportError:
        from oracle import scae_model
        from torch_scae_b200 import factory
        cfg = factory.prepare_model_params(**params)
        sd = {k: v.detach().clone().to(device).requires_grad_(v.is_floating_point())
              for k, v in factory.make_scae(params).state_dict().items()}
        opt = torch.optim.RMSprop([v for v in sd.values() if v.requires_grad], lr=3e-5, momentum=0.9,
                                  eps=1e-2 / float(batch) ** 2)

        def step():
            opt.zero_grad(set_to_none=True)
            res = scae_model.scae_forward(sd, cfg, image, None, training=True)
            loss, _ = scae_model.scae_loss(res, cfg, image, label)
            loss.backward()
            opt.step()
        return step, 'port', None


def cpu_train_throughput(config, batch, steps, warmup, threads=None):
    """images/s of fwd + loss + bwd + RMSprop of the reference on the host CPU -> (images/s, ms/step, threads, kind, times)."""
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    step, kind, _ = reference_stepper(config, batch, 'cpu')
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return batch * len(times) / total, 1000.0 * total / len(times), threads, kind, times


def cpu_hot_path_regions(config, batch, reps=3):
    """the two regions the kernels replace, timed in isolation on the host CPU (reference code):
    (a) TemplateBasedImageDecoder.forward + pdf.log_prob, forward + backward (part_decoder.py:152-243, distributions.py:41-48);
    (b) CapsuleObjectDecoder.forward, forward + backward, with its per-capsule MLPs timed separately so that the post-MLP
        region (object_decoder.py:160-236, :257-372) is the difference.  Returns None when only the port is available."""
    step, kind, handle = reference_stepper(config, batch, 'cpu')
    if kind != 'reference':
        return None
    model, image, label = handle
    with torch.no_grad():
        enc = model.part_encoder(image)
        templates = model.template_generator(feature=enc.feature, batch_size=batch).templates
        parts = torch.cat([enc.pose, 1. - enc.presence.unsqueeze(-1), enc.feature,
                           templates.view(*templates.shape[:2], -1)], -1)
        obj_encoding = model.obj_encoder(parts, enc.presence)
    leaf = lambda t: t.detach().clone().requires_grad_(True)

    def region_a():
        t, p, pr = leaf(templates), leaf(enc.pose), leaf(enc.presence)
        rec = model.part_decoder(templates=t, pose=p, presence=pr)
        rec.pdf.log_prob(image).sum().backward()

    def region_b():
        h = leaf(obj_encoding)
        res = model.obj_decoder(h, enc.pose.detach(), enc.presence.detach())
        (res.log_prob + res.posterior_mixing_prob.sum() + res.caps_presence.sum()).backward()

    layer = model.obj_decoder.capsule_layer

    def region_b_mlps():
        h = leaf(obj_encoding)
        raw = torch.stack([layer.mlps[i](h[:, i]) for i in range(layer.n_caps)], 1)
        ext = torch.cat([raw, torch.ones(batch, layer.n_caps, 1)], -1)
        torch.stack([layer.caps_mlps[i](ext[:, i]) for i in range(layer.n_caps)], 1).sum().backward()

    def best(fn):
        fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return min(ts) * 1e3
    a_ms, b_ms, mlp_ms = best(region_a), best(region_b), best(region_b_mlps)
    return dict(batch=batch, template_decoder_log_prob_fwd_bwd_ms=round(a_ms, 1),
                object_decoder_fwd_bwd_ms=round(b_ms, 1), object_decoder_mlps_fwd_bwd_ms=round(mlp_ms, 1),
                object_decoder_post_mlp_fwd_bwd_ms=round(b_ms - mlp_ms, 1),
                note='min of %d runs after one warm-up; post-MLP = object decoder - its per-capsule MLPs' % reps)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    config = config_of(args)
    batch = args.batch                                   # the GPU arm's per-GPU batch: same config on both arms
    steps = min(args.steps if args.steps else 10, 20)    # bounded: ~3 s per 1024-image step on 8-16 host cores
    warmup = min(args.warmup if args.warmup is not None else 3, 3)
    value, ms, threads, kind, times = cpu_train_throughput(config, batch, steps, warmup)
    times = sorted(times)
    sample = (f'{steps} timed steps of batch {batch} after {warmup} warm-ups ({kind}: '
              + ('the unmodified reference from baseline/_ref through its own factory / SCAE.forward / SCAE.loss'
                 if kind == 'reference' else 'oracle port of the reference op sequence; baseline/_ref is absent')
              + f'); min {1e3 * times[0]:.0f} / median {1e3 * times[len(times) // 2]:.0f} ms per step')
    line = dict(metric=METRIC, value=round(value, 2), unit=UNIT, impl='reference', n_gpus=args.gpus, steps=steps,
                warmup=warmup, ms_per_step=round(ms, 3), higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='fp32', data='synthetic',
                config=dict(workload=workload_name(config, batch), n_obj_caps=CONFIGS[config][1]['O'],
                            global_batch=batch, parallelism='cpu', device=f'host CPU, {threads} threads'),
                cpu_baseline=dict(value=round(value, 2), unit=UNIT, cores=threads, kind=kind, sample=sample),
                e2e=dict(value=round(value, 2), unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), file=RESULT_OUT, flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
class StepHarness:
    """One model + flat gradient bucket + RMSprop + synthetic batch of a BASELINE config on this rank's GPU."""

    def __init__(self, config, B, dev, rank, scae_overrides=None):
        import copy
        from torch_scae_b200 import ddp, factory
        params = copy.deepcopy(CONFIGS[config][0])
        if scae_overrides:
            params['scae_params'] = dict(params.get('scae_params', {}), **scae_overrides)
        self.config, self.B, self.dev = config, B, dev
        torch.manual_seed(42)                                   # identical initial weights on every rank
        self.model = factory.make_scae(params).to(dev).train()
        ddp.broadcast_parameters(self.model)
        # parameters, gradients and optimizer state live in flat buffers: gradients are assigned (not accumulated) and
        # copied into the bucket by one multi-tensor launch, the all-reduce is one NCCL call on the bucket, and the
        # RMSprop update (the reference's optimizer and hyper-parameters, base_experiment.py:47-53) is one kernel
        self.bucket = ddp.FlatGradBucket(self.model, assign=True, flat_params=True)
        self.opt = ddp.FlatRMSprop(self.bucket, lr=3e-5, momentum=0.9, eps=1e-2 / float(B) ** 2)
        torch.manual_seed(42 + rank)                            # different synthetic shard per rank
        C, H, W = params['image_shape']
        self.host_image = torch.rand(B, C, H, W).pin_memory()
        self.host_label = torch.randint(0, 10, (B,)).pin_memory()
        self.image, self.label = self.host_image.to(dev), self.host_label.to(dev)
        self.graphed, self.graph_note = None, 'disabled (--no-graph)'

    def step(self, img=None, lab=None):
        img = self.image if img is None else img
        lab = self.label if lab is None else lab
        self.bucket.zero()
        res = self.model(img)
        loss, _ = self.model.loss(res, img, lab)
        loss.backward()
        self.bucket.collect()
        self.bucket.all_reduce_mean()
        self.opt.step()
        return loss

    def capture(self, world):
        """The whole step (zero, forward, loss, backward, optimizer; the all-reduce stays eager between two graphs when
        world > 1) as a CUDA graph: torch_scae_b200/graph.py."""
        from torch_scae_b200 import graph
        try:
            self.graphed = graph.GraphedTrainStep(self.model, self.opt, self.bucket, self.image, self.label)
            self.graph_note = 'whole step captured' if world == 1 else f'whole step captured, {self.graphed.collective_note}'
        except Exception as exc:                            # noqa: BLE001 - report and fall back to eager launches
            self.graphed, self.graph_note = None, f'capture failed, eager launches: {type(exc).__name__}: {exc}'[:300]
            torch.cuda.synchronize()

    def run(self):
        return self.graphed() if self.graphed is not None else self.step()

    def release(self):
        self.graphed = None
        self.model = self.bucket = self.opt = None
        torch.cuda.empty_cache()


def run_gpu(args):
    import torch.distributed as dist
    from torch_scae_b200 import ops
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the fused path has no CPU fallback (use --impl reference for the CPU arm)')
    stray = [k for k in os.environ if k.startswith('SCAE_B200_') or k.startswith('SCAE_CAPS')]
    if stray:     # development switches (A/B timing, bisecting) must not shape a benchmark number
        raise SystemExit(f'bench.py: unset the development switches first: {stray}')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    # strict fp32 everywhere (no TF32 in the cuDNN / cuBLAS parts of the step)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = True

    config = config_of(args)
    shape = CONFIGS[config][1]
    B = args.batch
    steps = args.steps if args.steps else 50
    warmup = max(3, args.warmup if args.warmup is not None else 10)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(n):
            fn()
        end.record()
        barrier()
        ms = torch.tensor([start.elapsed_time(end)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / n

    h = StepHarness(config, B, dev, rank)
    for _ in range(warmup):
        h.step()
    assert h.bucket.check_views(), 'gradient views were replaced; flat bucket all-reduce would be stale'
    if not args.no_graph:
        h.capture(world)
    graphed = h.graphed
    host_loss = torch.zeros((), dtype=torch.float32).pin_memory()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # (1) device-resident throughput.  The per-entry-point kernel timing brackets every C-ABI call with CUDA events on
    # the launching stream, which only exists for eager launches: with a captured step the headline is the graph replay
    # and the kernel durations come from an eager pass over the same step right before it.
    eager_steps = steps if graphed is None else min(steps, 20)
    with ops.KernelTimer() as timer:
        ms_eager = timed(h.step, eager_steps)
    kstats, kmedian = timer.summary(), timer.medians()
    # >= 100 replays inside the timed region, whatever --steps says; `steps` is what the line reports
    reps = max(steps, 100) if graphed is not None else steps
    ms_step = timed(h.run, reps)

    # (2) end to end: pinned host batch -> device every step, loss read back to the host every step.
    # With the captured step the input pipeline is double buffered: the pinned host batch of step i+1 is copied to a
    # device staging buffer on a side stream while step i runs (GraphedTrainStep.stage), like a prefetching data
    # loader; every timed step still moves one full batch host->device and reads its loss back.
    def e2e_step():
        if graphed is not None:
            loss = graphed()                                # consumes the staged batch
            graphed.stage(h.host_image, h.host_label)       # next step's batch: H2D overlaps this step
        else:
            img = h.host_image.to(dev, non_blocking=True)
            lab = h.host_label.to(dev, non_blocking=True)
            loss = h.step(img, lab)
        host_loss.copy_(loss.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()           # the user reads the loss every step

    if graphed is not None:
        graphed.stage(h.host_image, h.host_label)           # batch of the first e2e step
    for _ in range(3):
        e2e_step()
    ms_e2e = timed(e2e_step, reps)
    clocks = sampler.summary() if rank == 0 else None
    graph_note = h.graph_note
    h2d = int(h.host_image.numel() * 4 + h.host_label.numel() * 8)
    h.release()
    graphed = None

    # (3) the other BASELINE configs, the class-default vote type and the strong-scaling point: short whole-step runs
    extra = {}

    def short_run(cfg, b, overrides=None, n=20):
        hx = StepHarness(cfg, b, dev, rank, overrides)
        for _ in range(5):
            hx.step()
        if not args.no_graph:
            hx.capture(world)
        for _ in range(3):
            hx.run()
        with ops.KernelTimer() as t2:
            fast_before = ops.caps_fast_path_count()
            hx.step()
            fast_calls = ops.caps_fast_path_count() - fast_before
        ms = timed(hx.run, n)
        out = dict(value=round(b * world * 1000.0 / ms, 1), unit=UNIT, ms_per_step=round(ms, 3), batch_per_gpu=b,
                   global_batch=b * world, cuda_graph=hx.graph_note.split(',')[0],
                   capsule_calls_on_the_persistent_path=f'{fast_calls} of 2',
                   kernels_ms={k: round(v[2] / v[0], 4) for k, v in t2.summary().items() if k.startswith('scae_tmpl_ll') or k.startswith('scae_caps_ll')})
        hx.release()
        return out

    if not args.no_extra:
        extra['configs'] = {}
        for cfg in ('mnist32', 'mnist10', 'stress', 'color'):
            if cfg == config:
                continue
            # colour: BASELINE configs[4] is batch 4096 on 8 GPUs = 512 per GPU
            b = 512 if cfg == 'color' else 1024
            extra['configs'][cfg] = dict(short_run(cfg, b), workload=CONFIGS[cfg][2])
        if config == 'mnist32':
            extra['vote_soft'] = dict(short_run('mnist32', B, dict(vote_type='soft')),
                                      note="SCAE's class default vote_type='soft' (stacked_capsule_auto_encoder.py:31): the "
                                           "decoder pose is the soft winner, so the capsule backward also receives "
                                           "g_soft_winner")
            if 8192 % world == 0:
                extra['strong_8192'] = dict(short_run('mnist32', 8192 // world, n=10),
                                            note='BASELINE configs[2] literally: global batch 8192 split over the ranks')

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    bytes_per_image = algorithmic_bytes(fused_color=True, **shape)   # SCAE.forward colours in-kernel
    kernels = {}
    plumbing = {}
    launches = 0
    for name, (calls, n_launch, total_ms) in kstats.items():
        # median over the calls of the eager pass: the event pair around a C-ABI call also spans what the host does
        # between the two records, and one late launch would move the mean of a 40 us kernel
        avg_ms = kmedian[name]
        launches += (n_launch // eager_steps) * steps      # launches of OUR kernels inside the timed region of `value`
        if name not in bytes_per_image:        # small kernels serving the callers of the hot paths (e.g. scae_colsum)
            plumbing[name] = dict(calls_per_step=calls // eager_steps, launches_per_step=n_launch // eager_steps,
                                  ms_per_step=round(total_ms / eager_steps, 4))
            continue
        gbs = bytes_per_image[name] * B / (avg_ms * 1e-3) / 1e9
        kernels[name] = dict(ms=round(avg_ms, 4), ms_is=f'median of {calls} calls', launches_per_call=n_launch // calls,
                             algorithmic_bytes=bytes_per_image[name] * B, achieved_gbs=round(gbs, 1),
                             frac_of_hbm_peak=round(gbs / peak, 4), share_of_step=round(avg_ms / ms_step, 4))
        if name in ISSUE_LANE_INSTR:
            t_issue = issue_roof_ms(name, B, (clocks or {}).get('sm_mhz') or 1965, **shape)
            kernels[name].update(binding_roof='sm issue rate', issue_roof_ms=round(t_issue, 4),
                                 frac_of_issue_roof=round(t_issue / avg_ms, 3),
                                 issue_roof_basis=f'{ISSUE_LANE_INSTR[name]} algorithmic lane-instructions per (pixel, '
                                                  'component)')
            if config in ('mnist32', 'mnist10'):
                t_meas = measured_issue_ms(name, B, (clocks or {}).get('sm_mhz') or 1965)
                if t_meas is not None:      # ncu-counted instruction stream of this build: issue-pipe utilisation
                    kernels[name].update(measured_issue_ms=round(t_meas, 4),
                                         issue_pipe_utilisation=round(t_meas / avg_ms, 3))
        else:
            kernels[name].update(binding_roof='hbm')
    dominant = max(kernels, key=lambda k: kernels[k]['ms'])
    traffic = ncu_traffic().get(dominant)
    roofline = dict(kernel=dominant, bound='hbm', achieved=kernels[dominant]['achieved_gbs'], peak=peak, unit='GB/s',
                    frac=kernels[dominant]['frac_of_hbm_peak'], traffic=traffic, peak_source=peak_src,
                    binding_roof=kernels[dominant]['binding_roof'],
                    frac_of_binding_roof=kernels[dominant].get('frac_of_issue_roof', kernels[dominant]['frac_of_hbm_peak']),
                    issue_pipe_utilisation=kernels[dominant].get('issue_pipe_utilisation'),
                    hbm_bound_kernels={k: v['frac_of_hbm_peak'] for k, v in kernels.items() if v['binding_roof'] == 'hbm'},
                    note='path-1 kernels are fp32-issue / shared-memory bound by construction (the B x K x H x W tensor '
                         'is never written); the HBM-shaped kernels are the two capsule ones (hbm_bound_kernels: in-step '
                         'fraction at this batch, inputs partly L2-resident; profiles/ holds the L2-flushed B=8192 runs)')

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        # (4) the reference on this box's host cores, BASELINE configs[0] (batch 128), bounded sample; its two hot-path
        # regions in isolation; and the same reference moved to this GPU (the "before" number)
        v, ms_cpu, threads, kind, times = cpu_train_throughput(config, args.cpu_batch, 12, 3)
        times = sorted(times)
        cpu = dict(value=round(v, 2), unit=UNIT, cores=threads, kind=kind,
                   sample=f'12 timed steps of batch {args.cpu_batch} (fwd+loss+bwd+RMSprop) after 3 warm-ups, min '
                          f'{1e3 * times[0]:.0f} / median {1e3 * times[len(times) // 2]:.0f} ms per step; '
                          + ('the unmodified reference (baseline/_ref) through its own API' if kind == 'reference'
                             else 'oracle port of the reference op sequence (baseline/_ref absent)')
                          + ' on the host CPU',
                   cpu_model=_cpu_model(), hot_path_regions=cpu_hot_path_regions(config, args.cpu_batch))
        if not args.no_extra:
            try:
                step_ref, kind_ref, _ = reference_stepper(config, B, dev)
                for _ in range(3):
                    step_ref()
                ms_ref = timed(step_ref, 5)
                extra['gpu_before'] = dict(value=round(B * 1000.0 / ms_ref, 1), unit=UNIT, ms_per_step=round(ms_ref, 2),
                                           kind=kind_ref, batch=B,
                                           note='the reference moved to this GPU with .to(device), stock PyTorch CUDA ops, '
                                                'eager, strict fp32; contains its per-step host->device helper copies '
                                                '(SURVEY.md section 9), so it is a pessimistic "before"')
            except Exception as exc:                        # noqa: BLE001
                extra['gpu_before'] = dict(unavailable=f'{type(exc).__name__}: {exc}'[:200])

    value = B * world * 1000.0 / ms_step
    e2e_value = B * world * 1000.0 / ms_e2e
    line = dict(metric=METRIC, value=round(value, 1), unit=UNIT, n_gpus=world, steps=steps, warmup=warmup,
                ms_per_step=round(ms_step, 4), higher_is_better=True, scaling='weak', vs_baseline=None, dtype='fp32',
                data='synthetic',
                config=dict(workload=workload_name(config, B), n_obj_caps=shape['O'],
                            global_batch=B * world, parallelism=f'dp{world}', tf32=False, cuda_graph=graph_note,
                            ms_per_step_eager=round(ms_eager, 4), timed_replays=reps,
                            l2_policy='per-step working set (>190 MB of activations per 1024 images) exceeds the 126 MB L2',
                            note="BASELINE.json's parenthetical says 10 obj caps; the reference's mnist.yaml:4 says 32 "
                                 "(used here); the 10-capsule variant is under extra.configs.mnist10"),
                e2e=dict(value=round(e2e_value, 1), unit=UNIT, ms_per_step=round(ms_e2e, 4),
                         h2d_bytes_per_step=h2d * world, d2h_bytes_per_step=4 * world),
                gpu_launches=launches, roofline=roofline, kernels=kernels, plumbing_kernels=plumbing, cpu_baseline=cpu,
                clocks=clocks, extra=extra)
    print(json.dumps(line), file=RESULT_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def _cpu_model():
    try:
        with open('/proc/cpuinfo') as f:
            for ln in f:
                if ln.startswith('model name'):
                    return ln.split(':', 1)[1].strip()
    except OSError:
        pass
    return 'unknown'


RESULT_OUT = sys.stdout


def claim_stdout():
    """The contract is ONE JSON line on stdout.  NCCL prints its version banner to fd 1 on the GPU boxes, and other
    libraries may chat there too, so keep a private handle on the real stdout for the result line and point fd 1 at
    stderr for everything else."""
    global RESULT_OUT
    sys.stdout.flush()
    RESULT_OUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=None)
    ap.add_argument('--warmup', type=int, default=None)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=1024, help='images per GPU per step')
    ap.add_argument('--n-obj-caps', type=int, default=32)
    ap.add_argument('--config', default=None, choices=sorted(CONFIGS), help='BASELINE config of the headline number')
    ap.add_argument('--cpu-batch', type=int, default=128, help='batch of the bounded CPU sample (BASELINE configs[0])')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extra', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch the step eagerly instead of replaying a CUDA graph')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
