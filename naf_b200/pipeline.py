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
    def __init__(self, model, depth: int = 2, device=None, host_ring_bytes: int = 0):
        """`host_ring_bytes` > 0: results larger than this are read back through a pinned ring of two
        buffers of that size instead of one pinned buffer of the full result size (every byte still
        crosses to the host every step; the caller gets the ring, i.e. only the tail of the result)."""
        self.model = model
        self.depth = int(depth)
        self.host_ring_bytes = int(host_ring_bytes)
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
        self._full_h = None                   # pinned host buffer(s) of the un-reduced result
        self._i = 0

    def _slot_buffers(self, slot, image_h, feats_h):
        cur = self._slots[slot]
        if (cur is None or cur[0].shape != image_h.shape or cur[1].shape != feats_h.shape
                or cur[0].dtype != image_h.dtype or cur[1].dtype != feats_h.dtype):
            # allocated ON the copy-in stream: the caching allocator then never hands out a block whose
            # previous (compute-stream) user may still have kernels queued
            with torch.cuda.stream(self.s_in):
                cur = (torch.empty(image_h.shape, dtype=image_h.dtype, device=self.device),
                       torch.empty(feats_h.shape, dtype=feats_h.dtype, device=self.device))
            self._slots[slot] = cur
        return cur

    def _download_full(self, out, ev):
        """D2H of the whole result straight from the model's output tensor (no device-side copy).  The
        copies of consecutive steps are ordered on the copy-out stream, so ONE pinned host buffer is
        enough (it holds the latest result once the returned event has completed)."""
        flat = out.permute(0, 2, 3, 1).reshape(-1) if out.dim() == 4 else out.reshape(-1)   # pixel-major storage order
        if not flat.is_contiguous():
            flat = out.contiguous().view(-1)
        nbytes = flat.numel() * flat.element_size()
        ring = self.host_ring_bytes if 0 < self.host_ring_bytes < nbytes else 0
        want = (flat.dtype, ring // flat.element_size() if ring else flat.numel())
        if self._full_h is None or self._full_h[0] != want:
            n = 2 if ring else 1
            self._full_h = (want, [torch.empty(want[1], dtype=flat.dtype, pin_memory=True) for _ in range(n)])
        bufs = self._full_h[1]
        self.s_out.wait_event(ev)
        with torch.cuda.stream(self.s_out):
            flat.record_stream(self.s_out)
            if ring:
                step = want[1]
                for j, o in enumerate(range(0, flat.numel(), step)):
                    chunk = flat[o:o + step]
                    bufs[j & 1][:chunk.numel()].copy_(chunk, non_blocking=True)
            else:
                bufs[0].copy_(flat, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.s_out)
        return bufs[0], done

    @torch.no_grad()
    def step(self, image_h: torch.Tensor, feats_h: torch.Tensor, output_size,
             reduce: Optional[Callable[[torch.Tensor], torch.Tensor]] = None):
        """Enqueue one forward.  Returns (result_host, done_event): `result_host` is a pinned host
        tensor that holds `reduce(out)` -- or, without `reduce`, the whole result in its pixel-major
        storage order (B, Ho, Wo, C) flattened -- once `done_event` has completed.  Host buffers are
        reused (every `depth` steps with `reduce`, every step without): consume or copy them first."""
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
        if reduce is None:
            ev = torch.cuda.Event()
            ev.record(main)
            self._consumed[slot] = ev
            res_h, done = self._download_full(out, ev)
            self._downloaded[slot] = done
            return res_h, done
        res = reduce(out)
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

    One graph per (shapes, dtypes, output_size, conv precision class, weight version).  Inputs are
    copied into the graph's static buffers on every call; the returned tensor is the graph's static
    output buffer and is overwritten by the next call with the same shapes -- clone it to keep it.

    A captured graph holds raw device pointers, so every tensor the forward reads from a host-side
    cache (RoPE cos/sin tables and their block means, integer tap tables, packed conv weights) is kept
    alive by the graph's entry: later calls with other shapes replace those caches with NEW tensors and
    never free or overwrite the ones an older graph points at.  In-place weight updates
    (`load_state_dict`, an optimiser step) bump the parameters' versions, which are part of the key: the
    next call captures a fresh graph from the new weights and the stale entries are dropped.
    """

    def __init__(self, model, warmup: int = 2):
        self.model = model
        self.warmup = int(warmup)
        self._entries = {}

    def parameters(self):
        return self.model.parameters()

    def _weights_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.model.parameters()) + \
            tuple((b.data_ptr(), b._version) for b in self.model.buffers())

    def _pinned_cache_tensors(self):
        """Every cached device tensor the forward that was just captured may have read."""
        from . import ops
        keep = []
        for m in self.model.modules():
            for name in ("_tables", "_mean_tables"):
                t = getattr(m, name, None)
                if t is not None:
                    keep.append(t)
            for name in ("_naf_wtc", "_naf_wcl"):
                t = getattr(m, name, None)
                if t is not None:
                    keep.append(t[1])
        keep.append(list(ops._tap_cache.values()))
        return keep

    @torch.no_grad()
    def __call__(self, image: torch.Tensor, features: torch.Tensor, output_size):
        if not image.is_cuda or not features.is_cuda:
            raise RuntimeError("GraphedNAF needs CUDA tensors (no CPU fallback)")
        size = (int(output_size[0]), int(output_size[1]))
        wkey = self._weights_key()
        key = (tuple(image.shape), image.dtype, tuple(features.shape), features.dtype, size,
               bool(torch.backends.cudnn.allow_tf32), image.device.index, wkey)
        e = self._entries.get(key)
        if e is None:
            # graphs captured from older weights can never be replayed again: drop them
            for k in [k for k in self._entries if k[-1] != wkey]:
                del self._entries[k]
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
            e = (graph, img_s, ft_s, out_s, self._pinned_cache_tensors())
            self._entries[key] = e
        graph, img_s, ft_s, out_s = e[:4]
        img_s.copy_(image)
        ft_s.copy_(features)
        graph.replay()
        return out_s
