"""GPU: CUDA-graph replay of the training step (graph.GraphedTrainStep, Trainer(cuda_graph=True)) reproduces the eager
loop, and the device-side Adam step counter matches the host-side one."""
import copy
import json
import os
import time

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_captured_adam_step_matches_the_host_counter_variant():
    from mri_interpolation_b200 import _lib
    gen = torch.Generator(device=DEV).manual_seed(5)
    n = 4099
    p0 = torch.randn(n + 1, device=DEV, generator=gen)[:n]  # count not a multiple of 4: scalar tail
    state = [[p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)] for _ in range(2)]
    pads = [[torch.zeros(n + 4, device=DEV) for _ in range(4)] for _ in range(2)]  # 16-byte aligned arenas
    for k in range(2):
        pads[k][0][:n] = p0
    step_dev = torch.zeros(1, dtype=torch.int64, device=DEV)
    hyper = torch.zeros(2, device=DEV)
    for step in range(1, 8):
        g = torch.randn(n, device=DEV, generator=gen)
        for k in range(2):
            pads[k][1][:n] = g
        a, b = pads
        _lib.call("mri_adam_step", a[0].data_ptr(), a[1].data_ptr(), a[2].data_ptr(), a[3].data_ptr(), n, step, 5e-3, 0.9, 0.999,
                  1e-8, 0.0, 1.0, 1, _lib.stream())
        _lib.call("mri_adam_step_captured", b[0].data_ptr(), b[1].data_ptr(), b[2].data_ptr(), b[3].data_ptr(), n,
                  step_dev.data_ptr(), hyper.data_ptr(), 5e-3, 0.9, 0.999, 1e-8, 0.0, 1.0, 1, _lib.stream())
        assert int(step_dev) == step
        torch.testing.assert_close(b[0][:n], a[0][:n], rtol=1e-6, atol=1e-9)
        assert float(b[1].abs().max()) == 0.0  # fused gradient clear
    del state


def _hash_model():
    from mri_interpolation_b200 import models
    torch.manual_seed(1337)
    return models.HashMLP(dim_in=4, n_levels=16, n_features_per_level=2, log2_hashmap_size=12, base_resolution=16,
                          finest_resolution=200, dim_hidden=64, dim_out=1, n_layers=2, batch_norm=False, lr=5e-3)


def _siren_model():
    from mri_interpolation_b200 import models
    torch.manual_seed(1337)
    return models.SirenNet(dim_in=3, dim_hidden=64, dim_out=1, n_layers=3, lr=1e-3)


@pytest.mark.parametrize("make,dim", [(_hash_model, 4), (_siren_model, 3)])
def test_graphed_training_step_reproduces_the_eager_loop(make, dim):
    from mri_interpolation_b200.graph import GraphedTrainStep
    gen = torch.Generator(device=DEV).manual_seed(3)
    batches = [(torch.rand(4096, dim, device=DEV, generator=gen), torch.rand(4096, 1, device=DEV, generator=gen)) for _ in range(10)]
    eager, graphed = make().to(DEV), make().to(DEV)
    opt_e, opt_g = eager.configure_optimizers(), graphed.configure_optimizers()
    losses_e = []
    for b in batches:
        loss = eager.training_step(b, 0)
        loss.backward()
        opt_e.step()
        opt_e.zero_grad()
        losses_e.append(float(loss))
    step = GraphedTrainStep(graphed, opt_g, batches[0])
    # the warm-up inside the constructor was rolled back: nothing has been trained yet
    assert opt_g.step_count == 0
    for pe, pg in zip(make().to(DEV).parameters(), graphed.parameters()):
        assert torch.equal(pe, pg)
    losses_g = [float(step(b)) for b in batches]
    assert opt_g.step_count == len(batches) and int(opt_g._step_dev) == len(batches)
    for le, lg in zip(losses_e, losses_g):
        assert abs(le - lg) <= 1e-6 * max(abs(le), 1e-3)
    for (name, pe), (_, pg) in zip(eager.named_parameters(), graphed.named_parameters()):
        torch.testing.assert_close(pg, pe, rtol=1e-5, atol=1e-7, msg=name)
    # a batch of another shape falls back to the eager path and keeps the device counter in step
    short = (batches[0][0][:1000], batches[0][1][:1000])
    assert not step.matches(short)
    loss = graphed.training_step(short, 0)
    loss.backward()
    opt_g.step()
    opt_g.zero_grad()
    assert opt_g.step_count == len(batches) + 1 and int(opt_g._step_dev) == len(batches) + 1


def test_trainer_fit_with_cuda_graph_matches_eager_fit():
    from mri_interpolation_b200.datamodules import DeviceBatchLoader
    from mri_interpolation_b200.pl_compat import Trainer
    gen = torch.Generator(device=DEV).manual_seed(9)
    n, bs = 10000 * 12 + 777, 10000  # the reference's batch size (config/base.py:63); ragged last batch
    coords, pixels = torch.rand(n, 4, device=DEV, generator=gen), torch.rand(n, 1, device=DEV, generator=gen)
    results = {}
    for mode in (False, True):
        model = _hash_model()
        loader = DeviceBatchLoader(coords, pixels, bs, shuffle=True, device=DEV, seed=11)
        trainer = Trainer(accelerator="gpu", max_epochs=3, logger=False, enable_checkpointing=False, cuda_graph=mode)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        trainer.fit(model, loader)
        torch.cuda.synchronize()
        results[mode] = (copy.deepcopy(model.state_dict()), time.perf_counter() - t0, trainer.global_step, trainer.graphed_steps,
                         float(trainer.callback_metrics["train_loss"]))
    sd_e, t_e, steps_e, g_e, loss_e = results[False]
    sd_g, t_g, steps_g, g_g, loss_g = results[True]
    assert steps_e == steps_g == 3 * 13 and g_e == 0 and g_g == 3 * 12
    assert abs(loss_e - loss_g) <= 1e-5 * max(abs(loss_e), 1e-3)
    for k in sd_e:
        if sd_e[k].is_floating_point():
            torch.testing.assert_close(sd_g[k], sd_e[k], rtol=1e-5, atol=1e-7, msg=k)


def test_graph_replay_is_faster_than_the_eager_step_at_the_reference_batch_sizes():
    """Reported (gpurun_out/graph_speed.json): ms/step of the eager loop vs graph replay at batch 4 096 and 10 000."""
    from mri_interpolation_b200 import models
    from mri_interpolation_b200.graph import GraphedTrainStep
    g4 = dict(n_levels=16, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16, finest_resolution=2489)
    report = {}
    for bs in (4096, 10000):
        gen = torch.Generator(device=DEV).manual_seed(bs)
        batch = (torch.rand(bs, 4, device=DEV, generator=gen), torch.rand(bs, 1, device=DEV, generator=gen))
        torch.manual_seed(1337)
        model = models.HashMLP(dim_in=4, dim_hidden=64, dim_out=1, n_layers=2, batch_norm=False, lr=5e-3, **g4).to(DEV)
        opt = model.configure_optimizers()

        def eager():
            loss = model.training_step(batch, 0)
            loss.backward()
            opt.step()
            opt.zero_grad()

        def wall(fn, reps):
            for _ in range(10):
                fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / reps * 1e3

        t_eager = wall(eager, 200)
        step = GraphedTrainStep(model, opt, batch)
        t_graph = wall(lambda: step(batch), 200)
        report[str(bs)] = {"eager_ms_per_step": t_eager, "graph_ms_per_step": t_graph,
                           "eager_coords_per_s": bs / t_eager * 1e3, "graph_coords_per_s": bs / t_graph * 1e3}
        assert t_graph < t_eager
    if os.path.isdir("gpurun_out"):
        with open("gpurun_out/graph_speed.json", "w") as f:
            json.dump(report, f)
