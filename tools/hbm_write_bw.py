import torch
x = torch.empty(4 << 30, dtype=torch.uint8, device="cuda")
y = torch.empty(4 << 30, dtype=torch.uint8, device="cuda")
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: x.zero_()); print("memset 4 GiB: %.3f ms -> %.2f TB/s write" % (ms, 4.295 / ms))
ms = t(lambda: x.fill_(3)); print("fill 4 GiB: %.3f ms -> %.2f TB/s write" % (ms, 4.295 / ms))
ms = t(lambda: y.copy_(x)); print("copy 4 GiB: %.3f ms -> %.2f TB/s read+write" % (ms, 2 * 4.295 / ms))
ms = t(lambda: x.view(torch.int32).sum()); print("sum 4 GiB: %.3f ms -> %.2f TB/s read" % (ms, 4.295 / ms))
