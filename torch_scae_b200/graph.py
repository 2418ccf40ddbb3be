"""CUDA-graph capture of the SCAE train step.

After the two likelihood paths are fused, a train step is ~700 small launches (cuDNN/cuBLAS pieces of the encoders, the
elementwise tail of the losses, four fused likelihood launches).  ``GraphedTrainStep`` records

    bucket.zero() -> forward -> SCAE.loss -> backward [-> all-reduce] -> optimizer.step()

once into CUDA graphs and replays them from static input buffers, so the host only pays two or three graph launches
per step.  The fused kernels are launched through the C ABI on ``torch.cuda.current_stream()``, i.e. on the capturing
stream, and allocate nothing themselves (their outputs and workspaces come from the graph's private torch pool), so
they are captured like any other node.  Random draws (the presence noises) use torch's graph-safe Philox offsets and
differ on every replay, as in eager mode.

With more than one rank the gradient all-reduce of the flat bucket is captured INSIDE the one graph (NCCL collectives
are capturable; the host then launches a single graph per step and the collective starts the moment the backward's last
kernel ends instead of after a host round trip).  If the capture of the collective fails on this software stack the
step falls back to graph A (zero/forward/backward), an eager all-reduce, graph B (optimizer step).
"""
import torch


class GraphedTrainStep:
    """Callable ``step(image, label) -> loss`` (a static device scalar, overwritten by every call).

    ``model``      SCAE (or any module with ``forward(image)`` and ``loss(res, image, label) -> (loss, log)``)
    ``optimizer``  a torch optimizer constructed with ``capturable=True``
    ``bucket``     ddp.FlatGradBucket of ``model`` (its ``zero`` / ``all_reduce_mean`` are part of the step)
    ``image``/``label`` example inputs: their shapes/dtypes fix the static buffers
    """

    def __init__(self, model, optimizer, bucket, image, label, warmup=3):
        self.model, self.optimizer, self.bucket = model, optimizer, bucket
        self.static_image = image.clone()
        self.static_label = label.clone() if label is not None else None
        self.split = False                      # True: two graphs around an eager all-reduce (the fall-back)
        self.collective_note = 'single process' if bucket.world == 1 else ''
        self.graphs = []
        self.loss = None
        self._capture(warmup)

    # ---- the step in eager form (also what gets captured) -------------------------------------------------------------
    def _fwd_bwd(self):
        self.bucket.zero()
        res = self.model(self.static_image)
        loss, _ = self.model.loss(res, self.static_image, self.static_label)
        loss.backward()
        self.bucket.collect()
        return loss.detach()

    # ---- warm-up must not train: parameters and optimizer state are put back in place afterwards ----------------------
    def _optimizer_tensors(self):
        return [(st, k) for st in self.optimizer.state.values() for k, v in st.items() if torch.is_tensor(v)]

    def _snapshot(self):
        params = [p.detach().clone() for p in self.bucket.params]
        state = [(st, k, st[k].detach().clone()) for st, k in self._optimizer_tensors()]
        return params, state

    def _restore(self, snap):
        params, state = snap
        with torch.no_grad():
            for p, saved in zip(self.bucket.params, params):
                p.copy_(saved)
            if state:
                for st, k, saved in state:
                    st[k].copy_(saved)
            else:
                # the optimizer allocated its state during the warm-up (torch optimizers start every state tensor at
                # zero); zero it IN PLACE -- the captured optimizer step holds on to these very tensors
                for st, k in self._optimizer_tensors():
                    st[k].zero_()

    def _capture(self, warmup):
        # cuDNN autotuning, lazy library initialisation and the optimizer's state allocation must happen before the
        # capture: run the eager step a few times on a side stream (the recipe torch.cuda.graphs documents)
        snap = self._snapshot()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._fwd_bwd()
                self.bucket.all_reduce_mean()
                self.optimizer.step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if not self.bucket.check_views():
            raise RuntimeError('gradient views were replaced during warm-up; cannot capture the step')
        self._restore(snap)

        pool = torch.cuda.graph_pool_handle()
        if self.bucket.world > 1:
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool, capture_error_mode='thread_local'):
                    self.loss = self._fwd_bwd()
                    self.bucket.all_reduce_mean()
                    self.optimizer.step()
                self.graphs.append(g)
                self.collective_note = 'NCCL all-reduce captured inside the graph'
                return
            except Exception as exc:                      # noqa: BLE001 - this stack cannot capture the collective
                torch.cuda.synchronize()
                self.split = True
                self.collective_note = f'eager NCCL all-reduce between two graphs ({type(exc).__name__})'
                self._restore(snap)
        g_a = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_a, pool=pool, capture_error_mode='thread_local'):
            self.loss = self._fwd_bwd()
            if not self.split:
                self.optimizer.step()
        self.graphs.append(g_a)
        if self.split:
            g_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_b, pool=pool, capture_error_mode='thread_local'):
                self.optimizer.step()
            self.graphs.append(g_b)

    # ---- input prefetch: the host->device copy of the NEXT batch overlaps the current step -----------------------------
    def stage(self, image, label=None):
        """Starts copying a (pinned host) batch into one of two device staging buffers on a side stream and returns at
        once; the next ``__call__()`` without arguments consumes it (a 6 MB device-to-device copy in front of the graph).
        Staging batch i+1 right after launching step i hides the PCIe transfer behind the step."""
        if not hasattr(self, '_copy_stream'):
            self._copy_stream = torch.cuda.Stream()
            self._stage_img = [torch.empty_like(self.static_image) for _ in range(2)]
            self._stage_lab = [torch.empty_like(self.static_label) if self.static_label is not None else None
                               for _ in range(2)]
            self._ready = [torch.cuda.Event() for _ in range(2)]
            self._consumed = [torch.cuda.Event() for _ in range(2)]
            self._staged, self._stage_idx, self._has_label = [], 0, [False, False]
        if len(self._staged) == 2:
            raise RuntimeError('both staging buffers hold batches that no step has consumed yet')
        k = self._stage_idx
        self._copy_stream.wait_event(self._consumed[k])     # no-op until a step has read this buffer
        with torch.cuda.stream(self._copy_stream):
            self._stage_img[k].copy_(image, non_blocking=True)
            self._has_label[k] = label is not None and self._stage_lab[k] is not None
            if self._has_label[k]:
                self._stage_lab[k].copy_(label, non_blocking=True)
            self._ready[k].record(self._copy_stream)
        self._staged.append(k)
        self._stage_idx = k ^ 1

    def _consume_staged(self):
        k = self._staged.pop(0)
        cur = torch.cuda.current_stream()
        cur.wait_event(self._ready[k])
        self.static_image.copy_(self._stage_img[k], non_blocking=True)
        if self._has_label[k]:                           # a batch staged without labels keeps the previous ones
            self.static_label.copy_(self._stage_lab[k], non_blocking=True)
        self._consumed[k].record(cur)

    def __call__(self, image=None, label=None, non_blocking=True):
        """Copies the batch into the static buffers (host or device source) and replays the step."""
        if image is None and getattr(self, '_staged', None):
            self._consume_staged()
        if image is not None and image is not self.static_image:
            self.static_image.copy_(image, non_blocking=non_blocking)
        if label is not None and label is not self.static_label:
            self.static_label.copy_(label, non_blocking=non_blocking)
        self.graphs[0].replay()
        if self.split:
            self.bucket.all_reduce_mean()
            self.graphs[1].replay()
        return self.loss
