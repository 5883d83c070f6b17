import sys, os, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "masked-diffusion-model_b200"))
from tests.test_trainer_gpu import _b200_setup
a, model, ema, opt, sched, tr = _b200_setup("base", graph=False)
x0 = (torch.rand(8, 3, 32, 32) * 2 - 1).cuda()
torch.manual_seed(11); tr.Scheduler.adopt_torch_rng("cuda")
for i in range(2):
    tr._run_batch(i, (x0,), 0, 1, 0, None, None)
tr._set_input(x0)
from mdm_b200 import train_ops
def stage_fwd():
    a_ = tr.args
    ti = tr._draw_timeindex(8, x0.device)
    ts = torch.index_select(tr._timesteps_table(x0.device), 0, ti)
    n = tr.Scheduler.get_black_area_num_pixels_time(ts)
    return ts, tr.Scheduler.degrade_training(n, tr.input, mean_option=a_.mean_option, mean_area=a_.mean_area, want_degrade_mask=False)
def s1():
    stage_fwd()
def s2():
    ts, d = stage_fwd()
    with torch.no_grad():
        model(d[0], ts)
def s3():
    ts, d = stage_fwd()
    net = model(d[0], ts).sample
    train_ops.residual_mse(net, d[0], tr.input)
def s4():
    ts, d = stage_fwd()
    net = model(d[0], ts).sample
    loss, rec = train_ops.residual_mse(net, d[0], tr.input)
    loss.backward()
def s5():
    tr._optimizer_tail_device()
def s6():
    tr._forward_backward()
for mode in ("global", "thread_local", "relaxed"):
    for name, fn in (("degrade", s1), ("fwd", s2), ("loss", s3), ("bwd", s4), ("tail", s5), ("full", s6)):
        torch.cuda.synchronize()
        try:
            fn(); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode=mode):
                fn()
            g.replay(); torch.cuda.synchronize()
            print(mode, name, "OK", flush=True)
        except Exception as e:
            print(mode, name, "FAIL", repr(e)[:200], flush=True)
            try:
                torch.cuda.synchronize()
            except Exception as e2:
                print("sync fail", repr(e2)[:100])
