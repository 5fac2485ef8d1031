import os, sys, time, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_interpolation_b200 import models, sweep, nifti
from mri_interpolation_b200 import functional as Fn
dev = torch.device("cuda")
G4 = dict(n_levels=16, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16, finest_resolution=2489)
torch.manual_seed(1337)
model = models.HashMLP(dim_in=4, dim_hidden=64, dim_out=1, n_layers=2, lr=5e-3, batch_norm=False, **G4).to(dev)
opt = model.configure_optimizers()
vol = nifti.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "sample_ankle_dyn_mri.nii.gz")).get_fdata(np.float32)
pix = torch.from_numpy(vol).flatten(); pix = ((pix - pix.min()) / (pix.max() - pix.min())).to(dev)
sampler = Fn.VoxelSampler(pix, vol.shape)
shape = (352, 352, 6, 29)
def sweep_ms(reps=3):
    sweep.dense_sweep(model, shape); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): out = sweep.dense_sweep(model, shape)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, float(out.abs().max()), float(out.mean())
print("fresh model sweep ms", sweep_ms())
n = 1 << 19
for k in range(6):
    for i in range(60):
        idx = torch.randint(0, sampler.total, (n,), device=dev)
        x, y = sampler.batch(idx)
        l = model.training_step((x, y), i); l.backward(); opt.step(); opt.zero_grad()
    t = model.encoder.levels[15].embedding.weight
    print(f"after {(k+1)*60} steps: loss {float(l):.5f} sweep ms", sweep_ms(), "table absmax", float(t.abs().max()), "frac |t|<1e-30", float((t.abs() < 1e-30).float().mean()))
