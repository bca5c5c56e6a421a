"""HBM bandwidth by direction on this box (torch kernels, CUDA events, best of 10): write-only (fill), read-only (sum), copy."""
import torch
n = 1 << 30
a = torch.empty(n, dtype=torch.bfloat16, device="cuda")
b = torch.empty(n, dtype=torch.bfloat16, device="cuda")
def best(fn, bytes_):
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return bytes_ / min(ts) / 1e6
print("write-only GB/s", best(lambda: a.zero_(), 2 * n))
print("read-only  GB/s", best(lambda: a.view(torch.int16).max(), 2 * n))
print("copy       GB/s", best(lambda: b.copy_(a), 4 * n))
