#!/usr/bin/env python
"""Benchmark of the NAF cross-scale neighbourhood-attention forward on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one `NAF.forward(image, features, target_size)` over one batch of synthetic input
(BASELINE.json configs[1] "C2": DINOv2-B/14, C=768, 448 -> 896, K=7, batch 8 PER GPU; weak scaling).
Rank 0 prints ONE JSON line.  Keys beyond the base contract:

  value        whole NAF.forward (tensor-core conv encoder + RoPE/key-pool + attention kernels), inputs
               resident in HBM
  hot_path     the same metric for the part SURVEY.md 8(a) scopes (guidance map + features -> out)
  e2e          NAF.forward through the public host-to-host API from PINNED HOST buffers: H2D of image and
               features AND D2H of the WHOLE result (19.7 GB at C2) inside the timed region, every step.
               `e2e_sample` is the same loop reading back only one value per low-res cell (what a consumer
               that keeps the result on the GPU would copy): labelled, not the headline.
  roofline     attention kernel, timed live with CUDA events on its launching stream, rated on the bytes
               the timed launch REALLY moves (guidance map at the encoder resolution, read through
               replication factors); `frac_8d` rates it on SURVEY.md 8(d)'s formula (x counted at the target
               resolution); `roofline_rep1` is a second launch with x materialised at the target resolution
               (where both byte counts coincide)
  configs      whole-forward ms, attention ms and roofline fraction for the other BASELINE.json configs
               (C1, C3 [4 images per GPU: B=32 over 8 GPUs], C4, C5)
  other_paths  operator-level kernel timings of the paths the headline step does not exercise: the backward
               (tensor-core cell kernel) and the tap-table / ratio-1 forward (union-window tensor-core kernel)
  cpu_baseline the UNMODIFIED reference modules (oracle/_ref, NATTEN replaced by oracle/natten_stub.py) on
               the host cores, bounded sample (N=1, rank 0 only)
  gpu_eager_reference  the same unmodified modules + stand-in run through torch eager on the B200 itself, one
               full-size image (a same-device proxy for "the reference on a GPU"; NOT NATTEN)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (batch per GPU, C, guidance side, target side, low-res side, K)
    "C1": (1, 384, 224, 224, 16, 7),      # configs[0]: the reference's own CPU-runnable case
    "C2": (8, 768, 448, 896, 32, 7),      # configs[1]: the metric's configuration (one GPU)
    "C3": (4, 1024, 518, 1036, 37, 11),   # configs[2]: batch 32 sharded over 8 GPUs = 4 per GPU
    "C4": (16, 768, 336, 1344, 24, 7),    # configs[3]: batch 16
    "C5": (4, 768, 512, 2048, 32, 7),     # configs[4]: batch 4
    # the reference's own benchmark shape (test/test_utils.py:16-25): B=1, C=384, 448 -> 448, lr 28, K=9
    "REF": (1, 384, 448, 448, 28, 9),
}
D_GUIDE = 256
PRECISION = ("attention: fp32 class (split-fp16 operands, 3 tcgen05 passes, fp32 accumulation); conv encoder: "
             "tf32 class (fp16-rounded operands, fp32 accumulation), i.e. PyTorch's default allow_tf32=True "
             "that the reference runs with")


def ncu_traffic(kernel, workload):
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return int(json.load(fh)[f"{kernel}:{workload}"]["bytes"])
    except Exception:
        return None


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------ clock sampler
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown"}
    NOTE = {0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x1: "gpu_idle", 0x2: "app_clocks"}

    def __init__(self, index: int, period: float = 0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            # honour CUDA_VISIBLE_DEVICES remapping
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.handle, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                for bit, name in {**self.BAD, **self.NOTE}.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_min_mhz": s[0], "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------- reference on the host cores
def cpu_sample(workload):
    """Bounded sample of a workload for the host: ONE image with the same per-pixel work (same C, D, K,
    ratio r) at 1/16 of the area.  C1 (the reference's own CPU-runnable case) is small enough to run at
    its full size."""
    _, C, gi, to, lo, K = WORKLOADS[workload]
    if workload == "C1":
        return dict(B=1, C=C, guide=gi, target=to, low=lo, K=K, r=to // lo, full=True)
    r = to // lo
    lo_s = max(K, lo // 4)
    return dict(B=1, C=C, guide=max(1, gi * lo_s // lo), target=lo_s * r, low=lo_s, K=K, r=r, full=False)


def reference_cpu_model(K):
    """(model, kind): the UNMODIFIED reference NAF (oracle/_ref or /root/reference, NATTEN replaced by
    the stub) when it is available, else the oracle port driven by naf_b200's torch modules."""
    from oracle import reference_runner as R

    torch.manual_seed(0)
    if R.available():
        ns = R.load()
        return ns.NAF(kernel_size=K).eval(), "reference"
    import naf_b200

    return naf_b200.NAF(kernel_size=K).eval(), "port"


def cpu_step(model, kind, image, feats, target, K):
    with torch.no_grad():
        if kind == "reference":
            return model(image, feats, target)            # the reference's own NAF.forward
        from oracle import naf_oracle as O

        x = model.image_encoder.guidance(image, target)   # torch-CPU conv encoder + oracle port
        return O.naf_forward(x.contiguous(), feats, 4, 4, K)


def time_cpu_reference(workload, steps, warmup):
    """Mpix/s of the reference's CPU path on a bounded sample of `workload`, all host threads."""
    s = cpu_sample(workload)
    model, kind = reference_cpu_model(s["K"])
    g = torch.Generator(device="cpu").manual_seed(5)
    image = torch.randn(s["B"], 3, s["guide"], s["guide"], generator=g)
    feats = torch.randn(s["B"], s["C"], s["low"], s["low"], generator=g)
    tgt = (s["target"], s["target"])
    for _ in range(warmup):
        cpu_step(model, kind, image, feats, tgt, s["K"])
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(model, kind, image, feats, tgt, s["K"])
    dt = (time.perf_counter() - t0) / max(1, steps)
    mpix = s["B"] * s["target"] ** 2 / 1e6
    what = ("the unmodified reference modules (src/model/naf.py NAF.forward, copied byte for byte into oracle/_ref) "
            "with NATTEN's na2d_qk/na2d_av replaced by oracle/natten_stub.py (NATTEN is not installable offline)"
            if kind == "reference" else
            "conv encoder on torch-CPU + the oracle port (RoPE, key pool, windowed attention); oracle/_ref absent")
    size = "full size" if s["full"] else "1 image, same per-pixel work at 1/16 of the area"
    desc = (f"{workload} {size}: guidance {s['guide']}^2 -> target {s['target']}^2, features "
            f"{s['C']}x{s['low']}x{s['low']}, r={s['r']}, K={s['K']}, batch 1; {what}")
    return mpix / dt, dt * 1e3, desc, kind


def time_gpu_eager_reference(workload, dev, steps=3, warmup=1):
    """The same UNMODIFIED reference modules (oracle/_ref) + torch NATTEN stand-in, but on the B200 itself: cuDNN
    convolutions, ATen elementwise / gather kernels, the replicated K and V and the (B,n,Ho,Wo,K*K) attention tensor
    in HBM like the reference's NATTEN path keeps them.  NOT NATTEN (not installable offline): a same-device proxy for
    "the reference on a GPU", one full-size image of the workload.  None when oracle/_ref is absent."""
    from oracle import reference_runner as R

    if not R.available():
        return None
    _, C, gi, to, lo, K = WORKLOADS[workload]
    torch.manual_seed(0)
    model = R.load().NAF(kernel_size=K).eval().to(dev)
    g = torch.Generator(device="cpu").manual_seed(5)
    image = torch.randn(1, 3, gi, gi, generator=g).to(dev)
    feats = torch.randn(1, C, lo, lo, generator=g).to(dev)
    with torch.no_grad():
        for _ in range(warmup):
            out = model(image, feats, (to, to))
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = model(image, feats, (to, to))
        e1.record()
        torch.cuda.synchronize(dev)
    del out
    ms = e0.elapsed_time(e1) / steps
    return {"value": round(to * to / 1e6 / (ms / 1e3), 3), "unit": "Mpix/s", "ms_per_image": round(ms, 2),
            "kind": "reference modules on the same GPU through torch eager (NOT NATTEN)",
            "sample": f"{workload}, ONE image at full size (guidance {gi}^2 -> target {to}^2, features {C}x{lo}x{lo}, K={K}); "
                      "unmodified src/model/naf.py NAF.forward from oracle/_ref, NATTEN's na2d_qk / na2d_av replaced by "
                      "oracle/natten_stub.py (pure torch, 16 target rows per gather), cuDNN convolutions with PyTorch's "
                      "default TF32"}


def run_reference_arm(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    value, ms, desc, kind = time_cpu_reference(args.workload, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "upsampled Mpix/s", "value": round(value, 5), "unit": "Mpix/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args.workload, args.gpus, sample=desc),
        "cpu_baseline": {"value": round(value, 5), "unit": "Mpix/s", "cores": cores, "kind": kind,
                         "sample": desc},
        "e2e": {"value": round(value, 5), "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "fps": round(1e3 / ms, 4),
        "note": "CPU arm: the reference has no CPU kernels of its own for this path (NATTEN is CUDA-first and not "
                "installable offline), so its unmodified Python modules run over a pure-torch NATTEN stand-in on "
                "the host cores; each step is a bounded sample of the workload (see config.cpu_sample)",
    }
    if args.workload != "C1":
        # BASELINE.json configs[0]: the reference's own CPU-runnable case, at its full size
        v1, ms1, d1, k1 = time_cpu_reference("C1", max(1, min(args.steps, 3)), 1)
        line["configs"] = {"C1": {"value": round(v1, 5), "unit": "Mpix/s", "ms_per_step": round(ms1, 2),
                                  "kind": k1, "cores": cores, "sample": d1}}
    print(json.dumps(line), flush=True)


def workload_config(name, n_gpus, sample=None):
    B, C, gi, to, lo, K = WORKLOADS[name]
    cfg = {"workload": f"{name}: NAF.forward, guidance {gi}x{gi} -> target {to}x{to}, features "
                       f"C={C} {lo}x{lo}, K={K}, D=256, 4 heads, batch {B}/GPU",
           "batch_per_gpu": B, "global_batch": B * n_gpus, "C": C, "guide": gi, "target": to,
           "low": lo, "K": K, "parallelism": f"batch-sharded x{n_gpus} (no data-path collective)",
           "l2": "working set per step (x 1.6 GB + out 19.7 GB at C2) >> 126 MB L2: no flush needed",
           "precision": PRECISION}
    if sample is not None:
        cfg["cpu_sample"] = sample
    return cfg


# ------------------------------------------------------------------------------------ GPU arm
class XattnTimer:
    """Wraps the one function that enqueues the attention kernel (naf_b200.ops._launch_xattn) with CUDA
    events recorded on the launching stream.  Lives here: the product carries no instrumentation."""

    def __init__(self, ops):
        self.ops, self.events, self.orig = ops, [], None

    def __enter__(self):
        self.orig = self.ops._launch_xattn

        def timed(p, dev):
            st = torch.cuda.current_stream(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            self.orig(p, dev)
            e1.record(st)
            self.events.append((e0, e1))

        self.ops._launch_xattn = timed
        return self

    def __exit__(self, *exc):
        self.ops._launch_xattn = self.orig

    def mean_ms(self):
        return sum(a.elapsed_time(b) for a, b in self.events) / max(1, len(self.events))


def xattn_bytes(B, C, to, lo, src):
    """(real, algorithmic-8d) bytes of one attention launch: real = what the launch moves when the
    guidance map is stored at `src` x `src` (read once), K and V maps read once, out written once;
    8d = SURVEY.md 8(d): 4*B*(D*Ho*Wo + C*h*w + C*Ho*Wo)."""
    real = 4.0 * B * (D_GUIDE * src * src + (D_GUIDE + C) * lo * lo + C * to * to)
    algo = 4.0 * B * (D_GUIDE * to * to + C * lo * lo + C * to * to)
    return real, algo


def other_paths(dev, peak):
    """Operator-level timings (kernels only, tensors resident, CUDA events, N = 1) of the paths the headline step
    does not exercise: the backward and the tap-table / ratio-1 forward (8f-3).  `frac` rates
    each launch on its algorithmic bytes: forward 4*(D + C) per pixel, backward 4*(C + 2*D) per pixel (dout and
    q read once, dq written once; windows and their gradients are < 1 %)."""
    import naf_b200
    from naf_b200 import _lib, ops

    def time_op(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n

    def entry(ms, nbytes, kernel):
        gbs = nbytes / (ms / 1e3) / 1e9
        return {"ms": round(ms, 4), "kernel": kernel, "bytes": int(nbytes), "achieved": round(gbs, 1), "unit": "GB/s",
                "frac": round(gbs / peak, 4)}

    out = {}
    g = torch.Generator(device="cpu").manual_seed(7)
    D = D_GUIDE
    # ---- backward at the reference's own backward benchmark (test/backward_speed.py: B=1, 448 <- 28, K=9) and on
    # one C2 image (896 <- 32, K=7, C=768)
    for name, (B, C, to, lo, K) in {"bwd_ref_protocol_c384": (1, 384, 448, 28, 9), "bwd_c2_image": (1, 768, 896, 32, 7)}.items():
        # NCHW-shaped views over pixel-major storage, what the autograd functions hand over: no packing pass in the timed call
        q = torch.randn(B, to, to, D, generator=g).to(dev).permute(0, 3, 1, 2)
        k = torch.randn(B, lo, lo, D, generator=g).to(dev).permute(0, 3, 1, 2)
        v = torch.randn(B, lo, lo, C, generator=g).to(dev).permute(0, 3, 1, 2)
        dout = torch.randn(B, to, to, C, generator=g).to(dev).permute(0, 3, 1, 2)
        tabs = naf_b200.RoPE(D, num_heads=4, base=100.0, rescale_coords=2.0).eval().to(dev).axis_tables(to, to)
        n0 = ops.launch_count("xattn_bwd_cell_tc")
        ms_b = time_op(lambda: ops.xattn_bwd(q, k, v, dout, 4, K, rope_tables=tabs))
        kern = "xattn_bwd (cell_tc)" if ops.launch_count("xattn_bwd_cell_tc") > n0 else "xattn_bwd (simt)"
        ms_f = time_op(lambda: ops.xattn(q, k, v, 4, K, rope_tables=tabs))
        out[name] = {"shape": f"B={B} C={C} {to}x{to} <- {lo}x{lo} K={K}",
                     "backward": entry(ms_b, 4.0 * B * to * to * (C + 2 * D), kern),
                     "forward": entry(ms_f, 4.0 * B * to * to * (C + D), "xattn (" + ops.select_algo(q.shape, v.shape, 4, K) + ")"),
                     "host_overhead_note": "both through naf_b200.ops (Python + ctypes): ~0.08 ms of host time per call bounds the small shapes"}
        del q, k, v, dout
    # ---- tap-table / ratio-1 forward: the reference's training shape (utils/training.py:28-50) and its denoising
    # shape (denoising.py:209-213,436-451) on the union-window tensor-core kernel vs the warp-per-pixel kernel
    for name, (B, Dq, n, C, to, lo, K) in {"train_32_from_13": (8, 256, 4, 384, 32, 13, 9),
                                           "denoise_256_r1_k15": (2, 256, 1, 3, 256, 256, 15)}.items():
        q = torch.randn(B, Dq, to, to, generator=g).to(dev)
        k = torch.randn(B, Dq, lo, lo, generator=g).to(dev)
        v = torch.randn(B, C, lo, lo, generator=g).to(dev)
        tabs = naf_b200.RoPE(Dq, num_heads=n, base=100.0, rescale_coords=2.0).eval().to(dev).axis_tables(to, to)
        ms_u = time_op(lambda: ops.xattn(q, k, v, n, K, rope_tables=tabs))
        ms_g = time_op(lambda: ops.xattn(q, k, v, n, K, rope_tables=tabs, algo=_lib.ALGO_GENERIC), n=3)
        out[name] = {"shape": f"B={B} D={Dq} heads={n} C={C} {to}x{to} <- {lo}x{lo} K={K}",
                     "auto": entry(ms_u, 4.0 * B * to * to * (C + Dq), "xattn (" + ops.select_algo(q.shape, v.shape, n, K) + ")"),
                     "generic_ms": round(ms_g, 4)}
        del q, k, v
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--algo", default="auto", choices=["auto", "generic", "cell_simt", "cell_tcws", "cell_tma", "union_tc"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-d2h", default="full", choices=["full", "sample", "both"],
                    help="what the e2e step reads back: the whole result (default; `e2e`), one value per "
                         "low-res cell (`e2e_sample`), or both loops")
    ap.add_argument("--configs", default="C1,C3,C4,C5",
                    help="other BASELINE.json configs reported in the `configs` block ('' = none)")
    ap.add_argument("--config-steps", type=int, default=5)
    ap.add_argument("--no-other-paths", action="store_true",
                    help="skip the `other_paths` block (backward and tap-table kernels at operator level)")
    ap.add_argument("--graph", type=int, default=0,
                    help="1: also time NAF.forward replayed as a CUDA graph (naf_b200.GraphedNAF)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import naf_b200
    from naf_b200 import _lib, dist as ndist, ops

    rank, world, local = ndist.init_from_env("nccl" if "RANK" in os.environ else None)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    algo = {v: k for k, v in _lib.ALGO_NAMES.items()}[args.algo]
    peak, peak_src = measured_hbm_peak()

    def barrier():
        if world > 1:
            torch.distributed.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    def timed(fn, steps, finish=None):
        """Device time of `steps` calls of fn, max over ranks, barrier+sync on both sides."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()   # side streams (transfers) join the timed stream before the closing event
        e1.record()
        barrier()
        return ndist.max_over_ranks(e0.elapsed_time(e1), dev)

    models = {}

    def model_for(K):
        # rank 0's random init broadcast to every rank with ONE NCCL broadcast
        if K not in models:
            torch.manual_seed(1234 + rank)
            m = naf_b200.NAF(kernel_size=K).eval().to(dev)
            m.upsampler.algo = algo
            models[K] = (m, ndist.broadcast_module_(m, src=0))
        return models[K]

    def roofline_entry(name, B, C, to, lo, K, src, ms, tag=None):
        real, algo8d = xattn_bytes(B, C, to, lo, src)
        chosen = ops.select_algo((B, D_GUIDE, to, to), (B, C, lo, lo), 4, K) if args.algo == "auto" else args.algo
        traffic = ncu_traffic(chosen, name + (f":{tag}" if tag else ""))
        ach = real / (ms / 1e3) / 1e9
        return {"bound": "hbm", "kernel": f"xattn ({chosen})", "achieved": round(ach, 1), "peak": peak,
                "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": traffic,
                "bytes": int(real), "bytes_what": f"bytes the timed launch moves: guidance map {src}x{src} (read "
                f"through replication x{to // src}) + K and V maps read once + out written once",
                "achieved_8d": round(algo8d / (ms / 1e3) / 1e9, 1), "frac_8d": round(algo8d / (ms / 1e3) / 1e9 / peak, 4),
                "bytes_8d": int(algo8d), "kernel_ms": round(ms, 4), "peak_source": peak_src}

    def run_config(name, steps, warmup):
        """Whole forward + attention-kernel time of one workload, inputs resident."""
        B, C, gi, to, lo, K = WORKLOADS[name]
        model, _ = model_for(K)
        g = torch.Generator(device="cpu").manual_seed(100 + rank)
        image = torch.randn(B, 3, gi, gi, generator=g).to(dev)
        feats = torch.randn(B, C, lo, lo, generator=g).to(dev)
        sink = {}

        def step():
            sink.clear()   # the previous result goes back to the allocator first (C4: 88.8 GB per result)
            sink["out"] = model(image, feats, (to, to))

        with torch.no_grad():
            for _ in range(warmup):
                step()
            torch.cuda.synchronize(dev)
            with XattnTimer(ops) as xt:
                total_ms = timed(step, steps)
                xattn_ms = xt.mean_ms()
        sink.clear()
        ms = total_ms / steps
        mpix = B * to * to / 1e6 * world
        src = gi if to % gi == 0 else to
        entry = {"workload": workload_config(name, world)["workload"], "global_batch": B * world,
                 "value": round(mpix / (ms / 1e3), 2), "unit": "Mpix/s", "ms_per_step": round(ms, 4),
                 "fps": round(B * world / (ms / 1e3), 2), "steps": steps,
                 "roofline": roofline_entry(name, B, C, to, lo, K, src, xattn_ms)}
        if name == "C1":
            # the latency-bound case (one 224 x 224 image: 22 short launches): the same forward replayed as ONE CUDA
            # graph through the product's own naf_b200.GraphedNAF, i.e. without the per-launch host cost
            graphed = naf_b200.GraphedNAF(model)

            def step_graph():
                sink["out"] = graphed(image, feats, (to, to))

            with torch.no_grad():
                for _ in range(warmup + 1):
                    step_graph()
                torch.cuda.synchronize(dev)
                gms = timed(step_graph, steps) / steps
            sink.clear()
            entry["graph_replay"] = {"ms_per_step": round(gms, 4), "value": round(mpix / (gms / 1e3), 2), "unit": "Mpix/s",
                                     "api": "naf_b200.GraphedNAF (inputs copied into the graph's static buffers every call)"}
        return entry

    # ============================================================ the metric's workload, in full
    B, C, gi, to, lo, K = WORKLOADS[args.workload]
    model, n_bcast = model_for(K)
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    image_h = torch.randn(B, 3, gi, gi, generator=g).pin_memory()
    feats_h = torch.randn(B, C, lo, lo, generator=g).pin_memory()
    image = image_h.to(dev)
    feats = feats_h.to(dev)
    target = (to, to)
    r = to // lo
    src = gi if to % gi == 0 else to
    sink = {}

    def step_full():
        sink["out"] = model(image, feats, target)

    x_holder = {}

    def step_hot():
        sink["out"] = model.upsample_from_guidance(x_holder["x"], feats, rep=x_holder["rep"])

    graphed = naf_b200.GraphedNAF(model) if args.graph else None
    fwd = graphed if graphed is not None else model
    out_bytes = 4 * B * C * to * to

    with torch.no_grad():
        # ---- warm-up (builds tables, packed weights and the caching-allocator pools)
        for _ in range(args.warmup):
            step_full()
        torch.cuda.synchronize(dev)
        sink.clear()

        # ---- timed: whole NAF.forward, inputs resident in HBM
        launches0 = ops.launch_count()
        sampler = ClockSampler(local)
        sampler.start()
        with XattnTimer(ops) as xt:
            total_ms = timed(step_full, args.steps)
            xattn_ms = xt.mean_ms()
        clocks = sampler.finish()
        launches = ops.launch_count() - launches0
        sink.clear()

        # ---- timed: the replaced part only (x, features ready -> out ready)
        x_holder["x"], x_holder["rep"] = model.image_encoder.guidance_source(image, target)
        for _ in range(2):
            step_hot()
        hot_ms = timed(step_hot, args.steps) / args.steps
        sink.clear()

        # ---- timed: the attention launch with the guidance map materialised at the TARGET resolution
        # (the shape every reference eval caller produces; SURVEY.md 8d asks for both image shapes)
        rep1 = None
        if x_holder["rep"] != (1, 1):
            ry, rx = x_holder["rep"]
            x_full = x_holder["x"].repeat_interleave(ry, 2).repeat_interleave(rx, 3).contiguous(
                memory_format=torch.channels_last)

            def step_rep1():
                sink["out"] = model.upsample_from_guidance(x_full, feats, rep=(1, 1))

            for _ in range(2):
                step_rep1()
            with XattnTimer(ops) as xt1:
                timed(step_rep1, max(3, args.steps // 2))
                rep1 = roofline_entry(args.workload, B, C, to, lo, K, to, xt1.mean_ms(), tag="rep1")
            sink.clear()
            del x_full
        x_holder.clear()

        graph_ms = None
        if graphed is not None:
            def step_graph():
                sink["out"] = graphed(image, feats, target)
            for _ in range(3):
                step_graph()
            graph_ms = timed(step_graph, args.steps) / args.steps
            sink.clear()

        # ---- timed: end to end from pinned host buffers through the public host-to-host API
        def e2e_loop(reduce, host_ring_bytes=0):
            pipe = naf_b200.HostPipeline(fwd, depth=2, host_ring_bytes=host_ring_bytes)
            got = {}

            def step():
                got["res"] = pipe.step(image_h, feats_h, target, reduce=reduce)[0]

            for _ in range(2):
                step()
            pipe.drain()
            torch.cuda.synchronize(dev)
            ms = timed(step, args.steps, finish=pipe.drain) / args.steps
            nbytes = got["res"].numel() * got["res"].element_size() if reduce is not None else out_bytes
            del pipe
            return ms, nbytes

        e2e_full = e2e_sample = None
        if args.e2e_d2h in ("full", "both"):
            try:
                ms, nb = e2e_loop(None)
                how = "one pinned host buffer of the result size"
            except RuntimeError as exc:   # pinned allocation of the full result refused by the host: ring
                ms, nb = e2e_loop(None, host_ring_bytes=1 << 30)
                how = f"pinned ring of 2 x 1 GiB (full-size pinned buffer failed: {str(exc)[:80]})"
            e2e_full = (ms, nb, how)
        if args.e2e_d2h in ("sample", "both") or e2e_full is None:
            e2e_sample = e2e_loop(lambda out: out[:, :, r // 2::r, r // 2::r])
        sink.clear()
        torch.cuda.empty_cache()

    ms_per_step = total_ms / args.steps
    mpix_step = B * to * to / 1e6 * world          # whole job, all ranks
    value = mpix_step / (ms_per_step / 1e3)
    h2d = int(image_h.numel() * 4 + feats_h.numel() * 4)
    api = ("naf_b200.HostPipeline.step (pinned host in -> pinned host out; H2D / D2H on side streams overlap the "
           "neighbouring steps' kernels)" + (" over naf_b200.GraphedNAF" if graphed is not None else ""))

    def e2e_entry(ms, nbytes, what):
        return {"value": round(mpix_step / (ms / 1e3), 3), "unit": "Mpix/s", "ms": round(ms, 4),
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(nbytes), "d2h": what, "api": api,
                "d2h_gbs": round(nbytes / (ms / 1e3) / 1e9, 2)}

    line = {
        "metric": "upsampled Mpix/s", "value": round(value, 3), "unit": "Mpix/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args.workload, world),
        "fps": round(B * world / (ms_per_step / 1e3), 3),
        "hot_path": {"value": round(mpix_step / (hot_ms / 1e3), 3), "unit": "Mpix/s", "ms": round(hot_ms, 4),
                     "what": "encoder-resolution guidance map + features resident -> out: pack V, RoPE+key-pool, attention"},
        "gpu_launches": int(launches),
        "roofline": roofline_entry(args.workload, B, C, to, lo, K, src, xattn_ms),
        "clocks": clocks,
        "weights_broadcast_elems": int(n_bcast),
    }
    if rep1 is not None:
        line["roofline_rep1"] = rep1
    if e2e_full is not None:
        line["e2e"] = e2e_entry(e2e_full[0], e2e_full[1], f"the WHOLE result (B,Ho,Wo,C) fp32, every step; {e2e_full[2]}")
        if e2e_sample is not None:
            line["e2e_sample"] = e2e_entry(e2e_sample[0], e2e_sample[1],
                                           "NOT host-to-host: only the upsampled features at every cell centre (B,C,h,w) are read back")
    else:
        line["e2e"] = e2e_entry(e2e_sample[0], e2e_sample[1],
                                "SAMPLE ONLY (--e2e-d2h sample): the upsampled features at every cell centre (B,C,h,w); "
                                "the result itself stays on the GPU")
    if graph_ms:
        line["graph_replay"] = {"value": round(mpix_step / (graph_ms / 1e3), 3), "unit": "Mpix/s", "ms": round(graph_ms, 4)}

    # ============================================================ the other BASELINE.json configs
    names = [n for n in args.configs.split(",") if n and n != args.workload]
    if names:
        line["configs"] = {}
        for name in names:
            try:
                with torch.no_grad():
                    line["configs"][name] = run_config(name, args.config_steps, 3)
            except Exception as exc:   # e.g. out of memory for a big batch on a shared GPU: report, go on
                line["configs"][name] = {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}
            torch.cuda.empty_cache()
        if "C3" in line["configs"] and "error" not in line["configs"]["C3"]:
            line["configs"]["C3"]["note"] = (f"BASELINE.json configs[2] = batch 32 sharded over 8 GPUs = 4 images per GPU; "
                                             f"this run: {world} GPU(s), global batch {4 * world}")

    # ============================================================ the other kernels behind the same operator API
    if world == 1 and not args.no_other_paths:
        try:
            line["other_paths"] = other_paths(dev, peak)
        except Exception as exc:
            line["other_paths"] = {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}
        torch.cuda.empty_cache()

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        v, ms, desc, kind = time_cpu_reference(args.workload, steps=2, warmup=1)
        line["cpu_baseline"] = {"value": round(v, 5), "unit": "Mpix/s", "cores": cores, "kind": kind,
                                "sample": desc, "ms_per_sample": round(ms, 1)}
        if args.workload != "C1":
            v1, ms1, d1, k1 = time_cpu_reference("C1", steps=2, warmup=1)
            line["cpu_baseline"]["C1_full_size"] = {"value": round(v1, 5), "unit": "Mpix/s", "ms": round(ms1, 1),
                                                    "kind": k1, "sample": d1}
        # ... and on the B200 itself (SURVEY.md 8d "optional third line": reference modules + torch-eager stand-in)
        try:
            torch.cuda.empty_cache()
            line["gpu_eager_reference"] = time_gpu_eager_reference(args.workload, dev)
        except Exception as exc:
            line["gpu_eager_reference"] = {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
