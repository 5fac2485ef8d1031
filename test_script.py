"""Plain ReLU MLP fitted to the 2-D+t slice of the sample volume - the reference's ``test_script.py`` (a copy of
notebook cells 1-7, test_script.py:16-97) on the B200 backend.  It is a demo, not a test: it exercises the
LightningModule / Trainer.fit / Trainer.predict / loader surface the reference's callers rely on."""
import os
import sys

import numpy as np
import torch

from mri_interpolation_b200 import metrics, models, nifti
from mri_interpolation_b200.config import BaseConfig
from mri_interpolation_b200.datamodules import DeviceBatchLoader
from mri_interpolation_b200.pl_compat import pl

torch.manual_seed(1337)


def main(epochs: int = 3, batch_size: int = 5000, dim_hidden: int = 352, n_layers: int = 8, root: str = None):
    data = nifti.load(BaseConfig().image_path).get_fdata(np.float32)[:, :, 3, :]  # 2-D + t slice
    Y = torch.from_numpy(np.ascontiguousarray(data)).reshape(-1, 1)
    Y = Y / Y.max()
    axes = [torch.linspace(0, 1, s) for s in data.shape]
    X = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1).reshape(len(Y), len(data.shape))
    dev = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
    train_loader = DeviceBatchLoader(X, Y, batch_size, shuffle=True, device=dev)
    test_loader = DeviceBatchLoader(X, Y, batch_size, shuffle=False, device=dev)

    model = models.BaseMLP(dim_in=len(data.shape), dim_hidden=dim_hidden, dim_out=1, n_layers=n_layers, lr=1e-4)
    trainer = pl.Trainer(accelerator="gpu" if torch.cuda.is_available() else "cpu", max_epochs=epochs, precision=32,
                         default_root_dir=root)
    trainer.fit(model, train_loader)
    pred = torch.concat(trainer.predict(model, test_loader))
    im = pred.reshape(data.shape).detach().cpu().numpy()
    psnr = metrics.peak_signal_noise_ratio(Y.reshape(data.shape).numpy(), im)
    print(f"ReLU MLP {n_layers}x{dim_hidden} after {epochs} epochs: PSNR {psnr:.2f} dB, {trainer.fit_seconds:.1f} s fit, "
          f"{len(train_loader) * epochs * batch_size / trainer.fit_seconds / 1e6:.1f} Mcoord/s")
    return psnr


if __name__ == "__main__":
    main(epochs=int(sys.argv[1]) if len(sys.argv) > 1 else 3)
