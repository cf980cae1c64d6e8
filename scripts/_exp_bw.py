import torch, time
dev = torch.device("cuda")
y = torch.empty(1 << 30, device=dev)  # 4.3 GB
x = torch.empty(1 << 30, device=dev)
def timeit(fn, it=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
t = timeit(lambda: y.fill_(1.0)); print("fill 4.29GB: %.3f ms  %.0f GB/s" % (t, 4.295 / t * 1e3))
t = timeit(lambda: y.zero_()); print("zero 4.29GB: %.3f ms  %.0f GB/s" % (t, 4.295 / t * 1e3))
t = timeit(lambda: y.copy_(x)); print("copy 4.29GB: %.3f ms  %.0f GB/s (r+w)" % (t, 2 * 4.295 / t * 1e3))
t = timeit(lambda: x.sum()); print("read 4.29GB: %.3f ms  %.0f GB/s" % (t, 4.295 / t * 1e3))
