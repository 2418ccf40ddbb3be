"""CUDA-graph capture of the SCAE train step.

After the two likelihood paths are fused, a train step is ~700 small launches (cuDNN/cuBLAS pieces of the encoders, the
elementwise tail of the losses, four fused likelihood launches).  ``GraphedTrainStep`` records

    bucket.zero() -> forward -> SCAE.loss -> backward [-> all-reduce] -> optimizer.step()

once into CUDA graphs and replays them from static input buffers, so the host only pays two or three graph launches
per step.  The fused kernels are launched through the C ABI on ``torch.cuda.current_stream()``, i.e. on the capturing
stream, and allocate nothing themselves (their outputs and workspaces come from the graph's private torch pool), so
they are captured like any other node.  Random draws (the presence noises) use torch's graph-safe Philox offsets and
differ on every replay, as in eager mode.

With more than one rank the gradient all-reduce stays OUTSIDE the graphs (graph A: zero/forward/backward, eager NCCL
all-reduce of the flat bucket, graph B: optimizer step) -- the collective is one call either way and keeping it eager
avoids tying the NCCL communicator's lifetime to a captured graph.
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
        self.split = bucket.world > 1
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

    def __call__(self, image=None, label=None, non_blocking=True):
        """Copies the batch into the static buffers (host or device source) and replays the step."""
        if image is not None and image is not self.static_image:
            self.static_image.copy_(image, non_blocking=non_blocking)
        if label is not None and label is not self.static_label:
            self.static_label.copy_(label, non_blocking=non_blocking)
        self.graphs[0].replay()
        if self.split:
            self.bucket.all_reduce_mean()
            self.graphs[1].replay()
        return self.loss
