"""Generate the golden fixtures in tests/golden/*.npz by RUNNING THE REFERENCE ITSELF
(/root/reference/code, imported through oracle/ref_shims.py) on CPU under fixed seeds.

Run in the build container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

The fixtures pin (a) the CPU random stream and its transforms, (b) `Scheduler.degrade_*` and
`get_schedule_shift_time`, (c) `Sampler.sample`, (d) both trainers' `_run_batch`, for the
reference-valid flag combinations listed in SURVEY.md section 8c."""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.mdm_oracle import default_args  # noqa: E402
from oracle.ref_shims import import_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

DEGRADE_CASES = {
    "thr_lin_1ch_deg_img": dict(select_degrade_pixel="thresholding", ddpm_schedule="linear", degrade_channel="1-channel", mean_option="degraded_area", mean_area="image-wise"),
    "thr_exp_3ch_deg_chan": dict(select_degrade_pixel="thresholding", ddpm_schedule="exponential", degrade_channel="3-channel", mean_option="degraded_area", mean_area="channel-wise"),
    "thr_lin_1ch_const": dict(select_degrade_pixel="thresholding", ddpm_schedule="linear", degrade_channel="1-channel", mean_option="0.0"),
    "idx_log_deg_img": dict(select_degrade_pixel="indexing", ddpm_schedule="log", mean_option="degraded_area", mean_area="image-wise"),
    "idx_log_deg_chan": dict(select_degrade_pixel="indexing", ddpm_schedule="log", mean_option="degraded_area", mean_area="channel-wise"),
    "idx_log_nondeg": dict(select_degrade_pixel="indexing", ddpm_schedule="log", mean_option="non_degraded_area"),
    "idx_log_const0": dict(select_degrade_pixel="indexing", ddpm_schedule="log", mean_option=0),
    "idx_log_deg_img_c1": dict(select_degrade_pixel="indexing", ddpm_schedule="log", mean_option="degraded_area", mean_area="image-wise", in_channel=1, out_channel=1),
    "idx_log_deg_img_bf16": dict(select_degrade_pixel="indexing", ddpm_schedule="log", mean_option="degraded_area", mean_area="image-wise", weight_dtype="bf16"),
}
SHIFT_TYPES = ["1-d_constant", "3-d_constant", "noise_reduction", "noise_std_reduction", "noise_with_perturbation", "non_shift"]

SAMPLER_CASES = {
    "indep_mom_noise": dict(select_degrade_pixel="indexing", ddpm_schedule="log", sampling_mask_dependency="independent", momentum_adaptive="base_momentum", shift_type="noise_with_perturbation", sample_latent_shape="uniform"),
    "prev_base_1d": dict(select_degrade_pixel="indexing", ddpm_schedule="log", sampling_mask_dependency="dependent_prev", momentum_adaptive="base_sampling", shift_type="1-d_constant", sample_latent_shape="normal"),
    "dept_mom_thr": dict(select_degrade_pixel="thresholding", ddpm_schedule="linear", degrade_channel="1-channel", mean_option="0", sampling_mask_dependency="dependent_t", momentum_adaptive="base_momentum", shift_type="3-d_constant", sample_latent_shape="zero"),
    "indep_mom_thr3_chan": dict(select_degrade_pixel="thresholding", ddpm_schedule="exponential", degrade_channel="3-channel", mean_option="non_degraded_area", mean_area="channel-wise", sampling_mask_dependency="independent", momentum_adaptive="base_momentum", shift_type="non_shift", sample_latent_shape="uniform"),
}


def mk_args(**kw):
    kw = dict(kw)
    wd = kw.pop("weight_dtype", "fp32")
    a = default_args(**kw)
    a.weight_dtype = torch.bfloat16 if wd == "bf16" else torch.float32
    return a


class ToyModel:
    """Deterministic elementwise 'denoiser' that is bit-reproducible on CPU and GPU."""

    def __init__(self, T, device="cpu"):
        self.T = float(T)
        self.device = torch.device(device)

    def __call__(self, x, t):
        scale = (t.float() / self.T).view(-1, 1, 1, 1)
        return SimpleNamespace(sample=x.float() * scale * -0.25)


class TinyNet(torch.nn.Module):
    """Small trainable stand-in for the trainers' golden (conv + time scale)."""

    def __init__(self, C):
        super().__init__()
        g = torch.Generator().manual_seed(1)
        self.conv = torch.nn.Conv2d(C, C, 3, padding=1)
        with torch.no_grad():
            self.conv.weight.copy_(torch.randn(self.conv.weight.shape, generator=g) * 0.1)
            self.conv.bias.copy_(torch.randn(self.conv.bias.shape, generator=g) * 0.1)

    @property
    def device(self):
        return self.conv.weight.device

    def forward(self, x, t):
        return SimpleNamespace(sample=self.conv(x.float()) * (1.0 + t.float().view(-1, 1, 1, 1) / 100.0))


class FakeAccelerator:
    sync_gradients = True
    is_main_process = True
    is_local_main_process = True

    def accumulate(self, model):
        import contextlib
        return contextlib.nullcontext()

    def backward(self, loss):
        loss.backward()

    def clip_grad_norm_(self, params, max_norm):
        return torch.nn.utils.clip_grad_norm_(list(params), max_norm)

    def wait_for_everyone(self):
        pass


def pack_mask(m: torch.Tensor) -> np.ndarray:
    return np.packbits(m.reshape(-1).numpy().astype(np.uint8))


def gen_rng():
    out = {}
    for seed in (0, 1234):
        rs = np.random.RandomState(seed)
        out[f"raw_{seed}"] = rs._bit_generator.random_raw(2000).astype(np.uint32)
        torch.manual_seed(seed); out[f"rand_{seed}"] = torch.rand(700).numpy()
        torch.manual_seed(seed); out[f"uniform_{seed}"] = torch.FloatTensor(700).uniform_(-1, 1).numpy()
        torch.manual_seed(seed); out[f"randperm8_{seed}"] = torch.randperm(8).numpy()
        torch.manual_seed(seed); out[f"randperm1024_{seed}"] = torch.randperm(1024).numpy()
        torch.manual_seed(seed); out[f"randint394_{seed}"] = torch.randint(0, 394, (700,)).numpy()
        torch.manual_seed(seed); out[f"normal_{seed}"] = torch.FloatTensor(1280).normal_(0.5, 2.0).numpy()
        torch.manual_seed(seed)
        torch.rand(1000)
        out[f"state_after_1000_{seed}"] = torch.get_rng_state().numpy()
    np.savez_compressed(os.path.join(OUT, "rng_kat.npz"), **out)


def gen_degrade(ref):
    out = {}
    S, B, T = 16, 5, 100
    for name, cfg in DEGRADE_CASES.items():
        a = mk_args(data_size=S, ddpm_num_steps=T, **cfg)
        C = a.in_channel
        g = torch.Generator().manual_seed(11)
        x0 = (torch.rand(B, C, S, S, generator=g) * 2 - 1).to(a.weight_dtype)
        torch.manual_seed(3)
        R = ref.Scheduler(a)
        Tp = R.update_ddpm_num_steps(T)
        ts = torch.tensor([1, 2, Tp // 2, Tp - 1, Tp])
        n = R.get_black_area_num_pixels_time(ts)
        d_img, masks, d_mask, mean_mask = R.degrade_training(n, x0, a.mean_option, a.mean_area)
        s_img, s_masks, s_mean = R.degrade_independent_base_sampling(n, x0.float(), a.mean_option, a.mean_area)
        w_img = R.degrade_with_mask(x0.float(), s_masks, a.mean_option, a.mean_area)
        out[f"{name}/x0"] = x0.float().numpy()
        out[f"{name}/timesteps"] = ts.numpy()
        out[f"{name}/Tp"] = np.int64(Tp)
        out[f"{name}/ratio_list"] = torch.as_tensor(R.ratio_list).numpy()
        out[f"{name}/n"] = n.numpy()
        out[f"{name}/degrade_img"] = d_img.float().numpy()
        out[f"{name}/masks"] = pack_mask(masks)
        out[f"{name}/masks_shape"] = np.array(masks.shape)
        out[f"{name}/degrade_mask"] = d_mask.float().numpy()
        out[f"{name}/fill"] = mean_mask[:, :, 0, 0].float().numpy()
        out[f"{name}/s_img"] = s_img.numpy()
        out[f"{name}/s_masks"] = pack_mask(s_masks)
        out[f"{name}/w_img"] = w_img.numpy()
        out[f"{name}/state_after"] = torch.get_rng_state().numpy()
    # dependent two-threshold draw
    a = mk_args(data_size=S, ddpm_num_steps=T, select_degrade_pixel="thresholding", ddpm_schedule="linear",
                degrade_channel="1-channel", mean_option="degraded_area", mean_area="image-wise")
    g = torch.Generator().manual_seed(11)
    x0 = torch.rand(B, 3, S, S, generator=g) * 2 - 1
    torch.manual_seed(5)
    R = ref.Scheduler(a)
    R.update_ddpm_num_steps(T)
    ts = torch.tensor([50, 60, 70, 80, 100])
    n_t, n_n = R.get_black_area_num_pixels_time(ts), R.get_black_area_num_pixels_time(ts - 1)
    r = R.degrade_dependent_base_sampling(n_t, n_n, x0, "degraded_area", "image-wise")
    out["dep2/x0"] = x0.numpy(); out["dep2/timesteps"] = ts.numpy()
    out["dep2/img_t"] = r[0].numpy(); out["dep2/mask_t"] = pack_mask(r[1]); out["dep2/img_n"] = r[3].numpy(); out["dep2/mask_n"] = pack_mask(r[4])
    np.savez_compressed(os.path.join(OUT, "degrade.npz"), **out)


def gen_shift(ref):
    out = {}
    S, B, T = 16, 6, 100
    for st in SHIFT_TYPES:
        for wd in ("fp32", "bf16"):
            a = mk_args(data_size=S, ddpm_num_steps=T, select_degrade_pixel="thresholding", ddpm_schedule="linear",
                        shift_type=st, noise_mean=0.25, weight_dtype=wd)
            torch.manual_seed(9)
            R = ref.Scheduler(a)
            R.update_ddpm_num_steps(T)
            ts = torch.tensor([1., 7., 33., 64., 99., 100.])
            like = torch.zeros(B, 3, S, S)
            sh = R.get_schedule_shift_time(ts, like)
            out[f"{st}/{wd}/shift"] = sh.float().numpy()
            out[f"{st}/{wd}/state_after"] = torch.get_rng_state().numpy()
    # quirk q7: B == W -> per-column scaling
    a = mk_args(data_size=8, ddpm_num_steps=50, select_degrade_pixel="thresholding", ddpm_schedule="linear",
                shift_type="noise_with_perturbation")
    torch.manual_seed(9)
    R = ref.Scheduler(a)
    R.update_ddpm_num_steps(50)
    ts = torch.arange(1, 9).float() * 5
    out["q7/shift"] = R.get_schedule_shift_time(ts, torch.zeros(8, 3, 8, 8)).numpy()
    np.savez_compressed(os.path.join(OUT, "shift.npz"), **out)


def gen_sampler(ref_sched, ref_samp):
    out = {}
    S, N, T = 16, 4, 10
    for name, cfg in SAMPLER_CASES.items():
        a = mk_args(data_size=S, ddpm_num_steps=T, sample_num=N, **cfg)
        torch.manual_seed(21)
        R = ref_sched.Scheduler(a)
        Tp = R.update_ddpm_num_steps(T)
        a.updated_ddpm_num_steps = Tp
        ts = R.get_timesteps_epoch(0, 1)
        samp = ref_samp.Sampler(None, a, R, [None, None, None])
        s0, vis = samp.sample(ToyModel(Tp), ts)
        out[f"{name}/sample_0"] = s0.numpy()
        out[f"{name}/timesteps"] = np.array(ts)
        out[f"{name}/sample_t_list"] = vis[0].numpy()
        out[f"{name}/sample_0_list"] = vis[5].numpy()
        out[f"{name}/degraded_t_list"] = vis[8].numpy()
        out[f"{name}/degraded_next_t_list"] = vis[10].numpy()
        out[f"{name}/state_after"] = torch.get_rng_state().numpy()
    np.savez_compressed(os.path.join(OUT, "sampler.npz"), **out)


HISTORY_EXTRA = {1: "shift_list", 2: "shifted_list", 3: "mask_list", 4: "shifted_result_list", 6: "degraded_mask_list",
                 7: "degraded_mask_next_list", 9: "difference_list"}


def gen_sampler_history(ref_sched, ref_samp):
    """the seven history tensors of sampler.py:116-126 that sampler.npz does not hold (same cases, same seed);
    a separate file so that the round-1 fixtures stay byte-identical"""
    out = {}
    S, N, T = 16, 4, 10
    for name, cfg in SAMPLER_CASES.items():
        a = mk_args(data_size=S, ddpm_num_steps=T, sample_num=N, **cfg)
        torch.manual_seed(21)
        R = ref_sched.Scheduler(a)
        Tp = R.update_ddpm_num_steps(T)
        a.updated_ddpm_num_steps = Tp
        ts = R.get_timesteps_epoch(0, 1)
        s0, vis = ref_samp.Sampler(None, a, R, [None, None, None]).sample(ToyModel(Tp), ts)
        for k, nm in HISTORY_EXTRA.items():
            v = vis[k]
            if nm.startswith("degraded_mask"):
                assert set(np.unique(v.numpy())) <= {0.0, 1.0}
                out[f"{name}/{nm}"] = pack_mask(v)
                out[f"{name}/{nm}_shape"] = np.array(v.shape)
            else:
                out[f"{name}/{nm}"] = v.numpy().astype(np.float16 if nm == "never" else np.float32)
    np.savez_compressed(os.path.join(OUT, "sampler_hist.npz"), **out)


def gen_train(ref_base, ref_ms):
    out = {}
    S, B, T = 16, 6, 100
    for method, mod in (("base", ref_base), ("mean_shift", ref_ms)):
        a = mk_args(data_size=S, ddpm_num_steps=T, select_degrade_pixel="indexing", ddpm_schedule="log",
                    mean_option="degraded_area", mean_area="image-wise", shift_type="noise_with_perturbation", method=method)
        g = torch.Generator().manual_seed(31)
        x0 = torch.rand(B, 3, S, S, generator=g) * 2 - 1
        net = TinyNet(3)
        opt = torch.optim.SGD(net.parameters(), lr=0.0)   # lr 0: gradients stay inspectable
        tr = mod.Trainer.__new__(mod.Trainer)
        tr.args, tr.model, tr.ema_model, tr.optimizer = a, net, None, opt
        tr.lr_scheduler = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0)
        tr.lr_list, tr.accelerator, tr.global_step = [], FakeAccelerator(), 0
        a.use_ema = False
        sched_mod = sys.modules["_mdmref_scheduler"]
        tr.Scheduler = sched_mod.Scheduler(a)
        Tp = tr.Scheduler.update_ddpm_num_steps(T)
        a.updated_ddpm_num_steps = Tp
        tr.timesteps_used_epoch = tr.Scheduler.get_timesteps_epoch(0, 1)
        opt.zero_grad = lambda *args, **kw: None           # keep .grad for the fixture
        torch.manual_seed(41)
        r = tr._run_batch(0, (x0,), 0, 1, 0, None, None)
        loss = r[0] if isinstance(r, tuple) else r
        out[f"{method}/x0"] = x0.numpy()
        out[f"{method}/loss"] = np.float64(loss)
        out[f"{method}/grad_w"] = net.conv.weight.grad.numpy()
        out[f"{method}/grad_b"] = net.conv.bias.grad.numpy()
        out[f"{method}/degraded"] = tr.degraded_img.float().numpy()
        out[f"{method}/recon"] = (tr.reconstructed_img if method == "base" else tr.inverse_shift_reconstructed_img).detach().float().numpy()
        out[f"{method}/state_after"] = torch.get_rng_state().numpy()
    np.savez_compressed(os.path.join(OUT, "train_step.npz"), **out)


if __name__ == "__main__":
    ref_sched, ref_samp, ref_base, ref_ms = import_reference("scheduler", "sampler", "trainer_masked", "trainer_masked_mean_shift")
    gen_rng()
    gen_degrade(ref_sched)
    gen_shift(ref_sched)
    gen_sampler(ref_sched, ref_samp)
    gen_sampler_history(ref_sched, ref_samp)
    gen_train(ref_base, ref_ms)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))
