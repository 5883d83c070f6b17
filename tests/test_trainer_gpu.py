"""GPU: the trainer drop-ins (`trainer_masked.Trainer`, `trainer_masked_mean_shift.Trainer`).

1. against the fixture produced by the REFERENCE's own `_run_batch` (tests/golden/train_step.npz,
   generator: tests/golden/make_golden.py:gen_train) with the same small conv net, seed and batch:
   degraded image bit-exact (masks from the device mt19937 stream), loss / gradients within the
   stated tolerance, CPU generator state identical after the step;
2. with the B200 denoiser + fused optimiser: the CUDA-graph step equals the eager step, the loss
   goes down on a fixed batch, EMA / LR bookkeeping matches the reference sequence."""
import copy

import numpy as np
import pytest
import torch

from tests.golden.make_golden import FakeAccelerator, TinyNet, mk_args
from tests.helpers import torch_state_words

pytestmark = pytest.mark.gpu


def _trainer(method, a, net, opt, sched, acc, ema=None):
    import trainer_masked
    import trainer_masked_mean_shift
    if method == "base":
        return trainer_masked.Trainer(a, None, None, net, ema, opt, sched, acc)
    return trainer_masked_mean_shift.Trainer(a, None, None, [None, None, None], net, ema, opt, sched, acc)


@pytest.mark.parametrize("method", ["base", "mean_shift"])
def test_run_batch_matches_reference_golden(golden, method):
    g = golden("train_step")
    a = mk_args(data_size=16, ddpm_num_steps=100, select_degrade_pixel="indexing", ddpm_schedule="log",
                mean_option="degraded_area", mean_area="image-wise", shift_type="noise_with_perturbation", method=method)
    a.use_ema = False
    a.timeindex_rng = "cpu_stream"          # the fixture was produced on CPU: indices come from the CPU generator
    a.materialize_visuals = True
    net = TinyNet(3).cuda()
    opt = torch.optim.SGD(net.parameters(), lr=0.0)
    opt.zero_grad = lambda *args, **kw: None
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0)
    tr = _trainer(method, a, net, opt, sched, FakeAccelerator())
    tr.prepare_schedule()
    tr.timesteps_used_epoch = tr.Scheduler.get_timesteps_epoch(0, 1)
    x0 = torch.from_numpy(g[f"{method}/x0"]).cuda()
    torch.manual_seed(41)
    tr.Scheduler.adopt_torch_rng("cuda")
    r = tr._run_batch(0, (x0,), 0, 1, 0, None, None)
    tr.Scheduler.release_rng_to_torch()
    loss = r[0] if isinstance(r, tuple) else r
    assert (len(r) == 3) if method == "base" else isinstance(r, float)
    # masks are bit-exact; the composite differs only by the summation order of the masked mean
    np.testing.assert_allclose(tr.degraded_img.float().cpu().numpy(), g[f"{method}/degraded"], atol=2e-6, rtol=0)
    kept_ref = g[f"{method}/degraded"] == g[f"{method}/x0"]          # pixels the reference's mask kept
    kept_mine = tr.degraded_img.float().cpu().numpy() == g[f"{method}/x0"]
    assert np.array_equal(kept_ref, kept_mine)                        # identical masks
    tol = 1e-5 if method == "base" else 5e-5
    assert abs(loss - float(g[f"{method}/loss"])) <= tol
    recon = tr.reconstructed_img if method == "base" else tr.inverse_shift_reconstructed_img
    np.testing.assert_allclose(recon.detach().float().cpu().numpy(), g[f"{method}/recon"], atol=5e-5, rtol=0)
    np.testing.assert_allclose(net.conv.weight.grad.cpu().numpy(), g[f"{method}/grad_w"], atol=tol, rtol=0)
    np.testing.assert_allclose(net.conv.bias.grad.cpu().numpy(), g[f"{method}/grad_b"], atol=tol, rtol=0)
    gk, gp = torch_state_words(g[f"{method}/state_after"])
    key, pos = torch_state_words(torch.get_rng_state().numpy())
    assert pos == gp and np.array_equal(key, gk)


def _b200_setup(method, C=3, S=32, B=8, graph=True, seed=0, T=100, optim="adamw", lr=2e-4):
    from mdm_b200.denoiser import UNet2DModelB200, default_config
    from mdm_b200.runtime import Accelerator, EMAModel, FusedOptimizer, get_scheduler
    a = mk_args(data_size=S, in_channel=C, out_channel=C, ddpm_num_steps=T, select_degrade_pixel="indexing",
                ddpm_schedule="log", mean_option="degraded_area", mean_area="image-wise",
                shift_type="noise_with_perturbation" if C == 3 else "1-d_constant", method=method)
    a.use_ema = True
    a.cuda_graph = graph
    a.timeindex_rng = "cpu_stream"
    torch.manual_seed(seed)
    model = UNet2DModelB200(device="cuda", **default_config(C, S))
    model.reset_parameters(seed=seed)
    ema = EMAModel(model.parameters(), decay=0.9999, use_ema_warmup=True, inv_gamma=1.0, power=0.75)
    opt = FusedOptimizer(model, optim, lr=lr)
    sched = get_scheduler("cosine", opt, num_warmup_steps=2, num_training_steps=100)
    acc = Accelerator(mixed_precision="bf16")
    model, opt, sched = acc.prepare(model, opt, sched)
    tr = _trainer(method, a, model, opt, sched, acc, ema)
    tr.prepare_schedule()
    tr.timesteps_used_epoch = tr.Scheduler.get_timesteps_epoch(0, 1)
    return a, model, ema, opt, sched, tr


@pytest.mark.parametrize("method", ["base", "mean_shift"])
def test_graph_step_equals_eager_step(method):
    g = torch.Generator().manual_seed(5)
    x0 = (torch.rand(8, 3, 32, 32, generator=g) * 2 - 1).cuda()
    losses = {}
    params = {}
    for graph in (False, True):
        # SGD: the parameter update is linear in the gradient, so the comparison is not amplified by Adam's
        # sign-like early steps on near-zero gradients (fp32 atomics reorder sums between runs)
        a, model, ema, opt, sched, tr = _b200_setup(method, graph=graph, optim="sgd", lr=1e-2)
        init = model.flat_param.clone()
        torch.manual_seed(11)
        tr.Scheduler.adopt_torch_rng("cuda")
        ls = []
        for i in range(6):           # 2 eager warm-up steps, capture, replays
            r = tr._run_batch(i, (x0,), 0, 1, 0, None, None)
            ls.append(r[0] if isinstance(r, tuple) else r)
        losses[graph] = ls
        params[graph] = (model.flat_param - init, ema.flat - init)
        assert ema.optimization_step == 6 and opt.step_count == 6 and tr.global_step == 6
        assert sched.last_epoch == 6
    # identical inputs, masks (same CPU-generator stream) and kernels; fp32 atomics in wgrad / GroupNorm
    # gradient reductions reorder sums, so equality is to rounding, not bitwise
    np.testing.assert_allclose(losses[True], losses[False], rtol=2e-3)
    rel = ((params[True][0] - params[False][0]).norm() / params[False][0].norm()).item()
    assert rel < 2e-2, rel          # relative difference of the accumulated UPDATE (6 steps)
    rel_ema = ((params[True][1] - params[False][1]).norm() / params[False][1].norm()).item()
    assert rel_ema < 2e-2, rel_ema
    assert losses[True][-1] < losses[True][0], losses[True]        # a fixed batch is being fitted


def test_training_reduces_loss_c1_shape():
    """BASELINE c1 shape (64 x 1 x 32 x 32), base trainer: 24 steps on a fixed batch"""
    a, model, ema, opt, sched, tr = _b200_setup("base", C=1, S=32, B=64, graph=True, T=10)
    g = torch.Generator().manual_seed(6)
    x0 = (torch.rand(64, 1, 32, 32, generator=g) * 2 - 1).cuda()
    torch.manual_seed(0)
    tr.Scheduler.adopt_torch_rng("cuda")
    ls = [tr._run_batch(i, (x0,), 0, 1, 0, None, None)[0] for i in range(24)]
    assert all(np.isfinite(ls)), ls
    # the timestep (hence the loss scale) is redrawn every step: compare averages
    assert np.mean(ls[-6:]) < 0.8 * np.mean(ls[:6]), ls


def test_ema_sample_then_graph_replay_uses_training_weights():
    """`_ema_sample()` = store, copy_to (EMA weights), Sampler.sample, restore.  The captured step
    graph contains no fp32 -> bf16 cast, so the restore must re-cast the mirror itself: the step after an EMA sample
    has to equal the same step of a run that never sampled."""
    g = torch.Generator().manual_seed(5)
    x0 = (torch.rand(8, 3, 32, 32, generator=g) * 2 - 1).cuda()
    out = {}
    for do_sample in (False, True):
        a, model, ema, opt, sched, tr = _b200_setup("base", graph=True, optim="sgd", lr=1e-2, T=20)
        a.sample_num, a.sample_latent_shape = 2, "uniform"
        torch.manual_seed(11)
        tr.Scheduler.adopt_torch_rng("cuda")
        for i in range(4):
            tr._run_batch(i, (x0,), 0, 1, 0, None, None)
        tr.Scheduler.release_rng_to_torch()
        state = torch.get_rng_state()
        if do_sample:
            ema.flat.mul_(0.0)                                   # make the EMA weights unmistakably different
            tr._ema_sample()
            assert torch.equal(model.flat_bf16, model.flat_param.to(torch.bfloat16))
        torch.set_rng_state(state)                               # the sampler consumed the CPU stream
        tr.Scheduler.adopt_torch_rng("cuda")
        out[do_sample] = tr._run_batch(4, (x0,), 0, 1, 0, None, None)[0]
        tr.Scheduler.release_rng_to_torch()
    assert abs(out[True] - out[False]) <= 2e-3 * abs(out[False]), out


def test_gradient_accumulation_follows_accelerate():
    """with gradient_accumulation_steps = 2 the optimiser, the LR schedule and zero_grad act on
    every second micro-batch only (accelerate's AcceleratedOptimizer / AcceleratedScheduler), both on the generic
    path (torch optimiser) and on the fused path."""
    from mdm_b200.runtime import Accelerator
    g = torch.Generator().manual_seed(5)
    x0 = (torch.rand(4, 3, 16, 16, generator=g) * 2 - 1).cuda()
    # generic path
    a = mk_args(data_size=16, ddpm_num_steps=100, select_degrade_pixel="indexing", ddpm_schedule="log",
                mean_option="degraded_area", mean_area="image-wise", method="base")
    a.use_ema, a.timeindex_rng = False, "cpu_stream"
    net = TinyNet(3).cuda()
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0 / (1 + s))
    acc = Accelerator(gradient_accumulation_steps=2)
    tr = _trainer("base", a, net, opt, sched, acc)
    tr.prepare_schedule()
    tr.timesteps_used_epoch = tr.Scheduler.get_timesteps_epoch(0, 1)
    torch.manual_seed(3)
    tr.Scheduler.adopt_torch_rng("cuda")
    w0 = net.conv.weight.detach().clone()
    tr._run_batch(0, (x0,), 0, 1, 0, None, None)
    assert torch.equal(net.conv.weight, w0) and net.conv.weight.grad.abs().sum() > 0        # accumulated, not applied
    assert sched.last_epoch == 0 and tr.global_step == 0
    g1 = net.conv.weight.grad.clone()
    tr._run_batch(1, (x0,), 0, 1, 0, None, None)
    assert not torch.equal(net.conv.weight, w0) and sched.last_epoch == 1 and tr.global_step == 1
    assert net.conv.weight.grad is None or float(net.conv.weight.grad.abs().sum()) == 0.0
    assert float(g1.abs().sum()) > 0
    # fused path
    a2, model, ema, opt2, sched2, tr2 = _b200_setup("base", graph=True, optim="sgd", lr=1e-2, T=20)
    tr2.accelerator.gradient_accumulation_steps = 2
    p0 = model.flat_param.clone()
    x1 = (torch.rand(8, 3, 32, 32, generator=g) * 2 - 1).cuda()
    tr2.Scheduler.adopt_torch_rng("cuda")
    tr2._run_batch(0, (x1,), 0, 1, 0, None, None)
    assert torch.equal(model.flat_param, p0) and sched2.last_epoch == 0 and opt2.step_count == 0
    assert float(model.flat_grad.abs().sum()) > 0 and ema.optimization_step == 0
    tr2._run_batch(1, (x1,), 0, 1, 0, None, None)
    assert sched2.last_epoch == 1 and opt2.step_count == 1          # (lr is still 0 at warm-up step 0: parameters unchanged)
    assert float(model.flat_grad.abs().sum()) == 0.0 and ema.optimization_step == 1 and tr2.global_step == 1
    tr2._run_batch(2, (x1,), 0, 1, 0, None, None)
    tr2._run_batch(3, (x1,), 0, 1, 0, None, None)
    assert not torch.equal(model.flat_param, p0) and sched2.last_epoch == 2 and opt2.step_count == 2
