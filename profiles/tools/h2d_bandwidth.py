import torch, time
n, m, full = 4096, 16000, 160000
pin = torch.empty((n, full), dtype=torch.int16).pin_memory()
dev = torch.empty((n, 179200), dtype=torch.int16, device='cuda')
stage = torch.empty((n, m), dtype=torch.int16, device='cuda')
cont = torch.empty((n, m), dtype=torch.int16).pin_memory()
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a=torch.cuda.Event(True); b=torch.cuda.Event(True); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/reps
bytes_ = n*m*2
ms = t(lambda: stage.copy_(cont, non_blocking=True)); print("contiguous pinned -> contiguous dev: %.3f ms  %.1f GB/s" % (ms, bytes_/ms/1e6))
ms = t(lambda: dev[:, 1600:1600+m].copy_(pin[:, 32000:32000+m], non_blocking=True)); print("strided pinned -> strided dev (2D): %.3f ms  %.1f GB/s" % (ms, bytes_/ms/1e6))
ms = t(lambda: stage.copy_(pin[:, 32000:32000+m], non_blocking=True)); print("strided pinned -> contiguous dev: %.3f ms  %.1f GB/s" % (ms, bytes_/ms/1e6))
big = torch.empty(n*full, dtype=torch.int16).pin_memory(); dbig = torch.empty(n*full, dtype=torch.int16, device='cuda')
ms = t(lambda: dbig.copy_(big, non_blocking=True), 3); print("1.3 GB contiguous: %.3f ms  %.1f GB/s" % (ms, n*full*2/ms/1e6))
