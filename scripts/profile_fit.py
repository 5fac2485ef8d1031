"""Host-side profile of Trainer.fit on the sample volume (cProfile, top cumulative entries) - where the launcher's loop spends
its time beyond the kernels.  usage (on a GPU box): python scripts/profile_fit.py"""
import cProfile, os, pstats, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mri_interpolation_b200 import config as cfgmod, datamodules, models
from mri_interpolation_b200.pl_compat import pl

dev = torch.device("cuda", 0)
cfg = cfgmod.HashConfig(); cfg.image_path, cfg.batch_size = bench.SAMPLE, 1 << 19
dm = datamodules.MriDataModule(config=cfg, device=dev); dm.prepare_data()
loader = dm.train_dataloader()
torch.manual_seed(1337)
model = models.HashMLP(dim_in=4, dim_hidden=64, dim_out=1, n_layers=2, lr=5e-3, batch_norm=False, **bench.G4)
pl.Trainer(accelerator="gpu", max_epochs=1, precision=32, enable_checkpointing=False, logger=False).fit(model, loader)
torch.cuda.synchronize()
tr = pl.Trainer(accelerator="gpu", max_epochs=4, precision=32, enable_checkpointing=False, logger=False)
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable(); tr.fit(model, loader); torch.cuda.synchronize(); pr.disable()
print("fit wall", time.perf_counter() - t0, "s for", 4 * len(loader), "steps")
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
