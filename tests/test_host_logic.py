"""CPU: host-side logic of the drop-in Scheduler (schedules, timestep subsets) against the oracle
and the reference goldens."""
import numpy as np
import pytest
import torch

import scheduler
from oracle.mdm_oracle import OracleRNG, OracleScheduler, default_args
from tests.golden.make_golden import DEGRADE_CASES, mk_args


@pytest.mark.parametrize("name", list(DEGRADE_CASES))
def test_schedule_tables(golden, name):
    g = golden("degrade")
    a = mk_args(data_size=16, ddpm_num_steps=100, **DEGRADE_CASES[name])
    S = scheduler.Scheduler(a)
    assert S.update_ddpm_num_steps(100) == int(g[f"{name}/Tp"])
    assert np.array_equal(torch.as_tensor(S.get_ratio_list()).numpy(), g[f"{name}/ratio_list"])
    n = S.get_black_area_num_pixels_time(torch.from_numpy(g[f"{name}/timesteps"]))
    assert np.array_equal(n.numpy(), g[f"{name}/n"])
    assert torch.equal(S.get_reverse_ratio_list(), torch.flip(torch.as_tensor(S.ratio_list), dims=(0,)))


@pytest.mark.parametrize("sched,S_,T,expect", [("log", 32, 1000, 394), ("log", 64, 1000, 802), ("log", 128, 1000, 1000),
                                               ("linear", 32, 1000, 1000), ("exponential", 32, 1000, 1000)])
def test_updated_num_steps(sched, S_, T, expect):
    a = default_args(data_size=S_, ddpm_num_steps=T, ddpm_schedule=sched,
                     select_degrade_pixel="indexing" if sched == "log" else "thresholding")
    assert scheduler.Scheduler(a).update_ddpm_num_steps(T) == expect
    assert OracleScheduler(a, OracleRNG(0)).update_ddpm_num_steps() == expect


@pytest.mark.parametrize("scale,epoch", [(1, 0), (3, 0), (3, 4), (3, 9)])
def test_timesteps_epoch(scale, epoch):
    a = default_args(data_size=16, ddpm_num_steps=64, scheduler_num_scale_timesteps=scale)
    S = scheduler.Scheduler(a)
    O = OracleScheduler(a, OracleRNG(0))
    S.update_ddpm_num_steps(64)
    O.update_ddpm_num_steps()
    assert S.get_timesteps_epoch(epoch, 10) == O.get_timesteps_epoch(epoch, 10)
    assert S.get_timesteps_epoch(epoch, 10)[-1] == 64


def test_reference_errors_kept():
    a = default_args(data_size=16, ddpm_num_steps=1000, ddpm_schedule="log")
    with pytest.raises(ValueError):
        scheduler.Scheduler(a).update_ddpm_num_steps(1000)      # more steps than pixels
    a = default_args(data_size=16, ddpm_num_steps=10, ddpm_schedule="sigmoid")
    with pytest.raises(TypeError):
        scheduler.Scheduler(a).update_ddpm_num_steps(10)        # quirk q5
    a = default_args(data_size=16, ddpm_num_steps=10, ddpm_schedule="nope")
    with pytest.raises(ValueError):
        scheduler.Scheduler(a).update_ddpm_num_steps(10)
