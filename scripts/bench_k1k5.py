"""K1 (degrade_training) / K5 (sampler step) / RNG micro-benchmark: achieved HBM GB/s vs algorithmic bytes.
    python scripts/bench_k1k5.py"""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "masked-diffusion-model_b200"))
import scheduler as sched_mod
from mdm_b200.config import default_args
from mdm_b200._lib import lib, ptr, check, stream_ptr
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, iters=10):
    for _ in range(2): fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3
res = []
for B, C, S, mode in ((128, 3, 32, "indexing"), (256, 3, 128, "thresholding"), (256, 3, 128, "indexing")):
    a = default_args(data_size=S, in_channel=C, out_channel=C, ddpm_num_steps=1000, ddpm_schedule="log" if mode == "indexing" else "linear",
                     select_degrade_pixel=mode, degrade_channel="1-channel", mean_option="degraded_area", mean_area="image-wise", sample_num=B)
    Sc = sched_mod.Scheduler(a); Tp = Sc.update_ddpm_num_steps(1000)
    torch.manual_seed(0); Sc.adopt_torch_rng("cuda")
    x0 = (torch.rand(B, C, S, S, device="cuda") * 2 - 1)
    ts = torch.randint(1, Tp + 1, (B,), device="cuda")
    n = Sc.get_black_area_num_pixels_time(ts)
    hw = S * S
    # masks (RNG, off the critical path) and K1 composite separately
    t_mask = timeit(lambda: Sc.make_mask_bytes(n, x0.device))
    mb = Sc.make_mask_bytes(n, x0.device)
    t_k1 = timeit(lambda: Sc._composite(x0, mb, 1, a.mean_option, a.mean_area, want_mask=True, want_degrade_mask=False))
    k1_bytes = B * (C * hw * 4 * 2 + hw * 4 + hw)          # read x0, write x_t, write 1-channel fp32 mask, read byte mask
    # K5: one sampler update
    net = torch.randn_like(x0); x_t = torch.randn_like(x0); shift = torch.randn_like(x0)
    mb2 = Sc.make_mask_bytes(n, x0.device).clone()
    x_next = torch.empty_like(x0); x_in_next = torch.empty_like(x0); s0 = torch.empty_like(x0)
    ws = torch.empty(max(1, 2 * lib().mdm_degrade_ws_floats(B, C, hw)), device="cuda")
    def k5():
        check(lib().mdm_sampler_step(ptr(x_t), ptr(net), ptr(shift), C * hw, hw, 1, ptr(mb), ptr(mb2), 1, 1, 0.0, 0, 1, 1,
                                     ptr(shift), C * hw, hw, 1, ptr(x_next), ptr(x_in_next), None, ptr(ws), B, C, hw, stream_ptr(x0.device)))   # (x0_hat is only materialised on the last iteration)
    t_k5 = timeit(k5)
    k5_bytes = B * C * hw * 4 * 5 + 2 * B * hw               # 5 fp32 image passes (+ the two byte masks)
    r = dict(shape=f"{B}x{C}x{S}x{S}", mask_mode=mode, mask_us=round(t_mask, 1), k1_us=round(t_k1, 1), k1_GBs=round(k1_bytes / t_k1 / 1e3, 1),
             k5_us=round(t_k5, 1), k5_GBs=round(k5_bytes / t_k5 / 1e3, 1))
    print(json.dumps(r), flush=True)
    res.append(r)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/k1k5.json", "w"))
