"""Pure-write / pure-read / copy bandwidth probes (context for the write-dominated kernels' roofline).
Measured on B200: fill 7247 GB/s, copy 6348 GB/s (read + write), torch.sum 5784 GB/s, cudaMemset 3857 GB/s."""
import torch, json
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize()
        best = min(best, s.elapsed_time(e))
    return best
N = 1 << 30
a = torch.empty(N, dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
af = a.view(torch.float32)
out = {}
out["memset_zero_GBs"] = N / timeit(lambda: a.zero_()) / 1e6
out["fill_f32_GBs"] = N / timeit(lambda: af.fill_(1.5)) / 1e6
out["copy_GBs_rw"] = 2 * N / timeit(lambda: b.copy_(a)) / 1e6
out["read_sum_GBs"] = N / timeit(lambda: af.sum()) / 1e6
# 1 read : 4 written, like the x4 up-sample (stock strided copy: far from the roofline)
q = a[: N // 4].view(torch.float32)
out["expand_1r4w_GBs"] = (N // 4 + N) / timeit(lambda: af.view(-1, 4).copy_(q.unsqueeze(1).expand(-1, 4))) / 1e6
print(json.dumps(out))
