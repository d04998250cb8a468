"""Host-to-host streaming front end of `NAF.forward`: pinned host inputs in, pinned host results out.

The reference is called with device tensors (`model(image.cuda(), features.cuda(), size)`,
evaluation/eval_seg_probing.py:100-105); a serving loop around it uploads every batch and reads a
result back.  `HostPipeline` does that with CUDA streams so that the transfers of step i+1 / i-1
overlap the kernels of step i:

    copy-in stream : H2D of image + features into one of `depth` device slots
    compute stream : NAF.forward (waits for its slot's upload), optional `reduce(out)` on device
    copy-out stream: D2H of the (reduced) result into a pinned host buffer

Every step still performs its own H2D and D2H; nothing is cached between steps.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch


class HostPipeline:
    def __init__(self, model, depth: int = 2, device=None):
        self.model = model
        self.depth = int(depth)
        # `model`: a NAF module, or any callable with its signature (e.g. GraphedNAF)
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("HostPipeline needs the model on a CUDA device (no CPU fallback)")
        self.s_in = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self._slots = [None] * self.depth     # (image_d, feats_d)
        self._uploaded = [torch.cuda.Event() for _ in range(self.depth)]
        self._consumed = [None] * self.depth  # compute finished reading the slot
        self._res_d = [None] * self.depth
        self._res_h = [None] * self.depth
        self._downloaded = [None] * self.depth
        self._i = 0

    def _slot_buffers(self, slot, image_h, feats_h):
        cur = self._slots[slot]
        if (cur is None or cur[0].shape != image_h.shape or cur[1].shape != feats_h.shape
                or cur[0].dtype != image_h.dtype or cur[1].dtype != feats_h.dtype):
            cur = (torch.empty(image_h.shape, dtype=image_h.dtype, device=self.device),
                   torch.empty(feats_h.shape, dtype=feats_h.dtype, device=self.device))
            self._slots[slot] = cur
        return cur

    @torch.no_grad()
    def step(self, image_h: torch.Tensor, feats_h: torch.Tensor, output_size,
             reduce: Optional[Callable[[torch.Tensor], torch.Tensor]] = None):
        """Enqueue one forward.  Returns (result_host, done_event): `result_host` is a pinned host
        tensor that holds `reduce(out)` (or `out`) once `done_event` has completed.  The host
        buffer of a slot is reused every `depth` steps: consume it (or copy it) before then."""
        main = torch.cuda.current_stream(self.device)
        slot = self._i % self.depth
        self._i += 1
        img_d, ft_d = self._slot_buffers(slot, image_h, feats_h)
        if self._consumed[slot] is not None:
            self.s_in.wait_event(self._consumed[slot])      # step i-depth no longer reads the slot
        with torch.cuda.stream(self.s_in):
            img_d.copy_(image_h, non_blocking=True)
            ft_d.copy_(feats_h, non_blocking=True)
            self._uploaded[slot].record(self.s_in)
        main.wait_event(self._uploaded[slot])
        if self._downloaded[slot] is not None:
            main.wait_event(self._downloaded[slot])          # the slot's result buffer has been drained
        out = self.model(img_d, ft_d, output_size)
        res = reduce(out) if reduce is not None else out
        if self._res_d[slot] is None or self._res_d[slot].shape != res.shape:
            self._res_d[slot] = torch.empty(res.shape, dtype=res.dtype, device=self.device)
            self._res_h[slot] = torch.empty(res.shape, dtype=res.dtype).pin_memory()
        self._res_d[slot].copy_(res)                         # gather / compact on the compute stream
        ev = torch.cuda.Event()
        ev.record(main)
        self._consumed[slot] = ev
        self.s_out.wait_event(ev)
        with torch.cuda.stream(self.s_out):
            self._res_h[slot].copy_(self._res_d[slot], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.s_out)
        self._downloaded[slot] = done
        return self._res_h[slot], done

    def drain(self):
        """Make the current stream wait for every outstanding transfer (end of a timed region)."""
        main = torch.cuda.current_stream(self.device)
        main.wait_stream(self.s_in)
        main.wait_stream(self.s_out)


class GraphedNAF:
    """CUDA-graph replay of `NAF.forward` for latency-bound shapes (the reference's own benchmark
    runs B=1: ~20 short kernels per forward, where launch overhead dominates).

        fast = naf_b200.GraphedNAF(model)
        out = fast(image, features, output_size)      # first call per shape: warm-up + capture

    One graph per (shapes, dtypes, output_size, conv precision class).  Inputs are copied into the
    graph's static buffers on every call; the returned tensor is the graph's static output buffer and
    is overwritten by the next call with the same shapes -- clone it to keep it.
    """

    def __init__(self, model, warmup: int = 2):
        self.model = model
        self.warmup = int(warmup)
        self._entries = {}

    def parameters(self):
        return self.model.parameters()

    @torch.no_grad()
    def __call__(self, image: torch.Tensor, features: torch.Tensor, output_size):
        if not image.is_cuda or not features.is_cuda:
            raise RuntimeError("GraphedNAF needs CUDA tensors (no CPU fallback)")
        size = (int(output_size[0]), int(output_size[1]))
        key = (tuple(image.shape), image.dtype, tuple(features.shape), features.dtype, size,
               bool(torch.backends.cudnn.allow_tf32), image.device.index)
        e = self._entries.get(key)
        if e is None:
            img_s, ft_s = image.clone(), features.clone()
            side = torch.cuda.Stream(image.device)
            side.wait_stream(torch.cuda.current_stream(image.device))
            with torch.cuda.stream(side):      # warm-up off the capture: builds tables, weight images, pools
                for _ in range(self.warmup):
                    self.model(img_s, ft_s, size)
            torch.cuda.current_stream(image.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out_s = self.model(img_s, ft_s, size)
            e = (graph, img_s, ft_s, out_s)
            self._entries[key] = e
        graph, img_s, ft_s, out_s = e
        img_s.copy_(image)
        ft_s.copy_(features)
        graph.replay()
        return out_s
