"""CPU, world_size 2 over gloo: the host-side data-parallel logic of `mdm_b200.runtime.Accelerator`
(what the reference gets from accelerate + torch DDP): gradient averaging equals the
single-process gradient on the concatenated batch, parameters are broadcast from rank 0, the DataLoader is
sharded batch-wise, the LR schedule advances world_size steps per optimiser step, every rank seeds 0 so the
CPU-generator streams (timesteps, masks) are identical across ranks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "masked-diffusion-model_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from mdm_b200.runtime import Accelerator, get_scheduler
    torch.manual_seed(0)                              # main_train_masked.py:441-445: every rank seeds 0
    acc = Accelerator(mixed_precision="no", device="cpu")
    assert acc.num_processes == world and acc.rank == rank and acc.is_main_process == (rank == 0)
    torch.manual_seed(100 + rank)                     # deliberately different initial weights per rank
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    sched = get_scheduler("linear", opt, num_warmup_steps=0, num_training_steps=100)
    g = torch.Generator().manual_seed(7)
    data = torch.randn(8, 4, 6, generator=g)          # 8 batches of 4
    target = torch.randn(8, 4, 3, generator=g)
    loader = [(data[i], target[i]) for i in range(8)]
    model, opt, loader, sched = acc.prepare(model, opt, loader, sched)
    w_after_prepare = [p.detach().clone() for p in model.parameters()]
    seen = []
    for x, y in loader:
        seen.append(x.clone())
    assert len(loader) == 4 and len(seen) == 4
    x, y = data[rank], target[rank]                   # rank r gets batches r, r+2, ...
    assert torch.equal(seen[0], x)
    with acc.accumulate(model):
        loss = torch.nn.functional.mse_loss(model(x), y)
        acc.backward(loss)
        acc.clip_grad_norm_(model.parameters(), 1e9)
        opt.step()
        sched.step()
    grads = [p.grad.detach().clone() for p in model.parameters()]
    # draws from the CPU generator after seeding 0 on every rank are identical across ranks
    torch.manual_seed(0)
    draw = torch.randint(0, 1000, (4,))
    out[rank] = dict(w=w_after_prepare, grads=grads, draw=draw, last_epoch=sched.last_epoch, lr=sched.get_last_lr()[0])
    acc.wait_for_everyone()
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_semantics_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    for a, b in zip(r0["w"], r1["w"]):
        assert torch.equal(a, b)                      # broadcast from rank 0
    for a, b in zip(r0["grads"], r1["grads"]):
        assert torch.allclose(a, b, atol=1e-7)        # all-reduced: same on both ranks
    assert torch.equal(r0["draw"], r1["draw"])
    assert r0["last_epoch"] == 2 and r1["last_epoch"] == 2      # accelerate: num_processes scheduler steps per step
    # reference: one process on the concatenated global batch (mean loss) -> same gradient
    torch.manual_seed(100)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    g = torch.Generator().manual_seed(7)
    data = torch.randn(8, 4, 6, generator=g)
    target = torch.randn(8, 4, 3, generator=g)
    x, y = torch.cat([data[0], data[1]]), torch.cat([target[0], target[1]])
    torch.nn.functional.mse_loss(model(x), y).backward()
    for got, p in zip(r0["grads"], model.parameters()):
        assert torch.allclose(got, p.grad, atol=1e-6)


def test_sampling_shards_are_independent_streams():
    """Sampling splits the batch per GPU with no communication: rank r is an independent
    reference process seeded seed0 + r; its oracle stream must not depend on world size."""
    from oracle.mdm_oracle import OracleRNG
    a = OracleRNG(5).raw(16)
    b = OracleRNG(6).raw(16)
    assert not (a == b).all()
    assert (OracleRNG(5).raw(16) == a).all()
