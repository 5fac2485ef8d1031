import os, sys, time, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_interpolation_b200 import models
dev = torch.device("cuda")
torch.manual_seed(1337)
net = models.SirenNet(dim_in=4, dim_hidden=256, n_layers=5, lr=1e-4).to(dev)
opt = net.configure_optimizers()
n = 1 << 18
xs = [torch.rand(n, 4, device=dev) * 2 - 1 for _ in range(4)]
ys = [torch.rand(n, 1, device=dev) for _ in range(4)]
def step(i, fresh):
    if fresh:
        x, y = torch.rand(n, 4, device=dev) * 2 - 1, torch.rand(n, 1, device=dev)
    else:
        x, y = xs[i % 4], ys[i % 4]
    l = net.training_step((x, y), i); l.backward(); opt.step(); opt.zero_grad()
    return l
for i in range(20): step(i, False)
torch.cuda.synchronize()
for label, fresh, sync_every in (("run-ahead reuse", False, 0), ("run-ahead fresh", True, 0), ("lag1 reuse", False, 1), ("run-ahead reuse again", False, 0)):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); a.record()
    prev = None
    for i in range(40):
        ev = torch.cuda.Event(); l = step(i, fresh); ev.record()
        if sync_every and prev is not None: prev.synchronize()
        prev = ev
    b.record(); t_cpu = time.perf_counter() - t0; torch.cuda.synchronize(); t_all = time.perf_counter() - t0
    print(f"{label:24s} gpu {a.elapsed_time(b)/40:.3f} ms/step  cpu-issue {t_cpu/40*1e3:.3f} ms/step  wall {t_all/40*1e3:.3f} ms/step  mem {torch.cuda.memory_reserved()/2**30:.1f} GiB")
