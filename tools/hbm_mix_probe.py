"""HBM bandwidth of this GPU by access mix: write-only (fill), read-only (sum), copy (read + write), each over 1 GiB,
best of 10 between CUDA events.  The row movers of the hot path are write-heavy (three_interpolate: 134 MB written for
43 MB read), so the copy figure of MEASURED_PEAKS.json is not their ceiling.  usage: python tools/hbm_mix_probe.py"""
import torch
dev = torch.device("cuda:0")
n = 1 << 28                                    # 1 GiB of fp32
a = torch.empty(n, dtype=torch.float32, device=dev).normal_()
b = torch.empty_like(a)
def best(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); t = min(t, e0.elapsed_time(e1))
    return t
gb = n * 4 / 1e9
print("write-only  fill_   %.1f GB/s" % (gb / best(lambda: b.fill_(1.0)) * 1e3))
print("write-only  zero_   %.1f GB/s" % (gb / best(lambda: b.zero_()) * 1e3))
print("read-only   sum     %.1f GB/s" % (gb / best(lambda: a.sum()) * 1e3))
print("copy        copy_   %.1f GB/s (read + write bytes)" % (2 * gb / best(lambda: b.copy_(a)) * 1e3))
c = torch.empty(n // 4, dtype=torch.float32, device=dev).normal_()
out = torch.empty(n // 4 * 3, dtype=torch.float32, device=dev).view(3, -1)
print("1 read : 3 written (broadcast copy) %.1f GB/s" % ((n // 4 * 4 * 4) / 1e9 / best(lambda: out.copy_(c.expand(3, -1))) * 1e3))
