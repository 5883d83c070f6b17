"""GPU: drop-in Sampler (side-stream mask/shift generation + fused K5 update) against the goldens
produced by the reference's own sampler.py and against the oracle."""
import numpy as np
import pytest
import torch

import sampler
import scheduler
from oracle.mdm_oracle import OracleRNG, OracleSampler, OracleScheduler
from tests.golden.make_golden import SAMPLER_CASES, ToyModel, mk_args
from tests.helpers import torch_state_words, unpack_mask

pytestmark = pytest.mark.gpu


def tol_for(a):
    # everything is exact fp32 algebra except (i) the masked mean (summation order) and (ii) normal_
    # noise (libm variants, 2.4e-7); both enter x_t once per step and are carried along 10 steps.
    return 1e-4


@pytest.mark.parametrize("name", list(SAMPLER_CASES))
@pytest.mark.parametrize("history", [False, True, "host"])
def test_sampler_against_reference_golden(golden, name, history, monkeypatch):
    g = golden("sampler")
    a = mk_args(data_size=16, ddpm_num_steps=10, sample_num=4, **SAMPLER_CASES[name])
    if history == "host":            # history tensors too large for the device budget: recorded straight into CPU tensors
        monkeypatch.setenv("MDM_HISTORY_DEVICE_BYTES", "0")
    a.sample_history = bool(history)
    S = scheduler.Scheduler(a)
    Tp = S.update_ddpm_num_steps(10)
    ts = S.get_timesteps_epoch(0, 1)
    torch.manual_seed(21)
    s0, vis = sampler.Sampler(None, a, S, [None, None, None]).sample(ToyModel(Tp, "cuda"), ts)
    assert s0.is_cuda and len(vis) == 11
    np.testing.assert_allclose(s0.cpu().numpy(), g[f"{name}/sample_0"], atol=tol_for(a), rtol=0)
    gk, gp = torch_state_words(g[f"{name}/state_after"])
    key, pos = torch_state_words(torch.get_rng_state().numpy())     # released back to torch's CPU generator
    assert pos == gp and np.array_equal(key, gk)
    if history:
        assert all(not v.is_cuda for v in vis)                     # handed back as CPU tensors, like the reference's
        np.testing.assert_allclose(vis[0].numpy(), g[f"{name}/sample_t_list"], atol=tol_for(a), rtol=0)
        np.testing.assert_allclose(vis[5].numpy(), g[f"{name}/sample_0_list"], atol=tol_for(a), rtol=0)
        np.testing.assert_allclose(vis[8].numpy(), g[f"{name}/degraded_t_list"], atol=tol_for(a), rtol=0)
        np.testing.assert_allclose(vis[10].numpy(), g[f"{name}/degraded_next_t_list"], atol=tol_for(a), rtol=0)
        # the other seven history tensors (sampler.py:116-126; fixture sampler_hist.npz): shift, shifted input, network
        # output, shifted result, both mask histories (bit-exact) and the difference incl. its stale last slot (q21)
        gh = golden("sampler_hist")
        for k, nm in ((1, "shift_list"), (2, "shifted_list"), (3, "mask_list"), (4, "shifted_result_list"), (9, "difference_list")):
            np.testing.assert_allclose(vis[k].numpy(), gh[f"{name}/{nm}"], atol=tol_for(a), rtol=0, err_msg=nm)
        for k, nm in ((6, "degraded_mask_list"), (7, "degraded_mask_next_list")):
            want = unpack_mask(gh[f"{name}/{nm}"], gh[f"{name}/{nm}_shape"])
            assert torch.equal(vis[k], want), nm
    else:
        assert all(v is None for v in vis)


def test_sampler_against_oracle_c1_shape():
    """BASELINE c1: 64 x 1 x 32 x 32, 10-step restoration."""
    a = mk_args(data_size=32, in_channel=1, out_channel=1, ddpm_num_steps=10, sample_num=64,
                select_degrade_pixel="indexing", ddpm_schedule="log", mean_option="degraded_area", mean_area="image-wise",
                shift_type="1-d_constant", sample_latent_shape="uniform")
    S = scheduler.Scheduler(a)
    Tp = S.update_ddpm_num_steps(10)
    ts = S.get_timesteps_epoch(0, 1)
    torch.manual_seed(6)
    s0, _ = sampler.Sampler(None, a, S, None).sample(ToyModel(Tp, "cuda"), ts)
    O = OracleScheduler(a, OracleRNG(6))
    O.update_ddpm_num_steps()
    r0, _ = OracleSampler(a, O, None).sample(ToyModel(Tp), ts)
    np.testing.assert_allclose(s0.cpu().numpy(), r0.numpy(), atol=1e-4, rtol=0)


def test_unsupported_modes_raise_like_reference():
    a = mk_args(data_size=16, ddpm_num_steps=10, sample_num=2, select_degrade_pixel="indexing", ddpm_schedule="log",
                momentum_adaptive="momentum")
    S = scheduler.Scheduler(a)
    S.update_ddpm_num_steps(10)
    with pytest.raises(UnboundLocalError):                                   # quirk q11
        sampler.Sampler(None, a, S, None).sample(ToyModel(10, "cuda"), [1, 2, 3])
