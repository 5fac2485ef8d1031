import os, sys, time, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_interpolation_b200 import models
from mri_interpolation_b200 import functional as Fn
dev = torch.device("cuda")
torch.manual_seed(1337)
net = models.SirenNet(dim_in=3, dim_hidden=1024, n_layers=8, lr=1e-4).to(dev)
opt = net.configure_optimizers()
n = 1 << 17
pix = torch.rand(512 * 512 * 512, device=dev) * 2 - 1
sampler = Fn.VoxelSampler(pix, (512, 512, 512), norm_siren=True)
index = torch.randint(0, sampler.total, (32, n), device=dev)
def step(i):
    x, y = sampler.batch(index[i % 32])
    l = net.training_step((x, y), i); l.backward(); opt.step(); opt.zero_grad()
    return l
evs = [torch.cuda.Event(enable_timing=True) for _ in range(101)]
cpu = []
evs[0].record()
inflight = []
for i in range(100):
    t0 = time.perf_counter()
    step(i)
    evs[i + 1].record()
    inflight.append(evs[i + 1])
    if len(inflight) > 2:
        inflight.pop(0).synchronize()
    cpu.append((time.perf_counter() - t0) * 1e3)
torch.cuda.synchronize()
gpu = [evs[i].elapsed_time(evs[i + 1]) for i in range(100)]
print("gpu ms per step:", " ".join(f"{g:.1f}" for g in gpu))
print("cpu ms per step:", " ".join(f"{c:.1f}" for c in cpu))
print("reserved GiB", torch.cuda.memory_reserved() / 2**30, "num cudaMalloc", torch.cuda.memory_stats()["num_device_alloc"])
