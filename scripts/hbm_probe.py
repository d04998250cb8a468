"""What HBM delivers for write-only / read-only / copy streams (torch kernels, CUDA events)."""
import torch
dev = torch.device("cuda", 0)
n = 5 * 1024 ** 3 // 4 * 4  # 5 Gi floats = 20 GiB? no: 5Gi elements * 4 B = 20 GiB
n = 4 * 1024 ** 3           # 4 Gi floats = 16 GiB
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n // 2, dtype=torch.float32, device=dev)
def timeit(fn, nbytes, name, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{name:28s} {nbytes / best / 1e6:8.0f} GB/s  ({best:.3f} ms)")
timeit(lambda: a.fill_(1.0), n * 4, "fill_ (write only) 16 GiB")
timeit(lambda: a.zero_(), n * 4, "zero_ (memset) 16 GiB")
timeit(lambda: torch.sum(a), n * 4, "sum (read only) 16 GiB")
timeit(lambda: b.copy_(a[: n // 2]), n * 4, "copy 8 GiB -> 8 GiB")
c = a[: n // 8]
timeit(lambda: torch.add(a[: n // 8 * 7], 1.0, out=a[: n // 8 * 7]), n // 8 * 7 * 8, "in-place add (r+w)")
