"""One small forward + backward of the B200 denoiser on cuda:0 checked against the fp32 oracle restatement
(oracle/unet_ref.py).  Used by `__graft_entry__.smoke()`; it lives under tests/ because it imports the oracle (the
product package never does)."""
from __future__ import annotations

import torch


def run(C=3, S=32, B=4):
    from oracle.unet_ref import UNet2DModelRef, unet_config   # test infrastructure (checker only)
    from mdm_b200.denoiser import UNet2DModelB200, default_config
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = UNet2DModelRef(**unet_config(C, S)).cuda()
    mine = UNet2DModelB200(device="cuda", **default_config(C, S))
    mine.load_state_dict(ref.state_dict())
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(B, C, S, S, device="cuda", generator=g) * 2 - 1
    x0 = torch.rand(B, C, S, S, device="cuda", generator=g) * 2 - 1
    t = torch.tensor([3.0, 77.0, 400.0, 999.0][:B], device="cuda")
    loss_ref = torch.nn.functional.mse_loss(x + ref(x, t).sample, x0)
    loss_ref.backward()
    mine.train()
    mine.zero_grad()
    out = mine(x, t).sample
    loss = torch.nn.functional.mse_loss(x + out, x0)
    loss.backward()
    rel = lambda a, b: ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()
    with torch.no_grad():
        e_out = rel(out, ref(x, t).sample)
    assert e_out <= 2e-2, f"denoiser output differs from the fp32 oracle: rel L2 {e_out:.3e}"
    assert abs(loss.item() - loss_ref.item()) <= 5e-3 * abs(loss_ref.item()), (loss.item(), loss_ref.item())
    got = mine.state_dict_grads()
    num = sum(((got[n].float() - p.grad.float()) ** 2).sum().item() for n, p in ref.named_parameters())
    den = sum((p.grad.float() ** 2).sum().item() for _, p in ref.named_parameters())
    e_grad = (num / den) ** 0.5
    assert e_grad <= 3e-2, f"denoiser gradients differ from the fp32 oracle: rel L2 {e_grad:.3e}"
    print(f"denoiser smoke ok: out rel L2 {e_out:.2e}, loss {loss.item():.5f} vs {loss_ref.item():.5f}, grad rel L2 {e_grad:.2e}")
