"""HBM bandwidth by access mix on this B200 (context for the write-dominated pyramid build): pure write (memset), pure read
(sum reduction), copy (read + write).  CUDA events, 2 GB buffers (>> L2), median of 10."""
import json, statistics, torch
n = 2 << 30
a = torch.empty(n, dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
af = a.view(torch.float32)
def t(fn, bytes_):
    ts = []
    for i in range(13):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(e0.elapsed_time(e1))
    return round(bytes_ / (statistics.median(ts) * 1e-3) / 1e9, 1)
res = {"write_only_memset_GBps": t(lambda: a.zero_(), n), "write_only_fill_f32_GBps": t(lambda: af.fill_(1.5), n),
       "read_only_sum_GBps": t(lambda: af.sum(), n), "copy_GBps_read_plus_write": t(lambda: b.copy_(a), 2 * n)}
print(json.dumps(res))
