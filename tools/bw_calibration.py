"""Calibration: pure-write and copy bandwidth of this GPU with torch kernels (context for the roofline denominator)."""
import json, torch
dev = torch.device("cuda:0")
n = 6_606_028_800 // 4  # the obs buffer of 8_arena at B=65536, in floats
x = torch.empty(n, device=dev); y = torch.empty(n, device=dev)
def timeit(f, reps=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t_fill = timeit(lambda: x.fill_(1.0))
t_zero = timeit(lambda: x.zero_())
t_copy = timeit(lambda: y.copy_(x))
print(json.dumps({"bytes": n * 4, "fill_ms": t_fill, "fill_GBps": n * 4 / t_fill / 1e6, "memset_ms": t_zero, "memset_GBps": n * 4 / t_zero / 1e6,
                  "copy_ms": t_copy, "copy_GBps_rw": 2 * n * 4 / t_copy / 1e6}))
