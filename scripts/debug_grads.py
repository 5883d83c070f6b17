import sys, os, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "masked-diffusion-model_b200"))
from oracle.unet_ref import UNet2DModelRef, unet_config
from mdm_b200.denoiser import UNet2DModelB200, default_config
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
C, S, B = 3, 32, 4
torch.manual_seed(3)
ref = UNet2DModelRef(**unet_config(C, S)).cuda()
mine = UNet2DModelB200(device="cuda", **default_config(C, S)); mine.load_state_dict(ref.state_dict())
g = torch.Generator(device="cuda").manual_seed(2)
x = torch.rand(B, C, S, S, device="cuda", generator=g) * 2 - 1
x0 = torch.rand(B, C, S, S, device="cuda", generator=g) * 2 - 1
t = torch.tensor([5.0, 100.0, 500.0, 900.0], device="cuda")
loss_ref = torch.nn.functional.mse_loss(x + ref(x, t).sample, x0); loss_ref.backward()
mine.train(); mine.zero_grad()
loss = torch.nn.functional.mse_loss(x + mine(x, t).sample, x0); loss.backward()
got = mine.state_dict_grads()
rows = []
for name, p in ref.named_parameters():
    a, b = got[name].float(), p.grad.float()
    rows.append(((a - b).norm().item() / (b.norm().item() + 1e-30), b.norm().item(), a.norm().item(), name))
order = [n for n, _ in ref.named_parameters()]
with open("gpurun_out/grads.txt", "w") as f:
    f.write(f"loss {loss.item()} ref {loss_ref.item()}\n")
    for r in rows:
        f.write("%.4e  ref_norm %.4e  mine_norm %.4e  %s\n" % r)
print("done")
