"""CPU: initial latent (sampler.py:46-83), data-mean histogram (main_train_masked.py:60-87) and loss weight
(scheduler.py:780-794) of the drop-ins against outputs of the REFERENCE itself (tests/golden/make_golden_latent.py).
Everything here is CPU-generator work in the reference, so the bar is bit-exactness, including the generator state
left behind (the next draws must match)."""
import os

import numpy as np
import pytest
import torch

import main_train_masked as M
import sampler
import scheduler
from oracle.mdm_oracle import default_args

Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "latent.npz"))


class _DS(torch.utils.data.Dataset):
    def __init__(self, data):
        self.data = data

    def __len__(self):
        return len(self.data)

    def __getitem__(self, i):
        return self.data[i], 0, 0


@pytest.mark.parametrize("area,nbin", [("image-wise", 12), ("channel-wise", 5)])
@pytest.mark.parametrize("shape", ["data", "zero", "normal", "uniform"])
def test_initial_latent_matches_reference(area, nbin, shape):
    a = default_args(data_size=8, in_channel=3, out_channel=3, sample_num=nbin, mean_area=area,
                     sample_latent_shape=shape, select_degrade_pixel="indexing", ddpm_schedule="log")
    a.weight_dtype = torch.float32
    data = torch.from_numpy(Z["data"])
    hist = M.compute_mean_histogram(_DS(data), a)
    key = f"{area}.{shape}"
    if shape == "data":
        assert np.array_equal(hist[2].numpy(), Z[f"{key}.cum"])
        for c, e in enumerate(hist[1]):
            assert np.array_equal(e.numpy(), Z[f"{key}.edges{c}"])
    smp = sampler.Sampler(None, a, scheduler.Scheduler(a), hist)
    torch.manual_seed(5)
    lat = smp._get_latent_initial(None)
    assert tuple(lat.shape) == (nbin, 3, 8, 8)
    assert np.array_equal(lat.contiguous().numpy(), Z[f"{key}.latent"])
    assert np.array_equal(torch.rand(4).numpy(), Z[f"{key}.next_rand"])      # same number of draws consumed


def test_grid_latent_raises_like_the_reference():
    a = default_args(data_size=8, sample_num=4, sample_latent_shape="grid", select_degrade_pixel="indexing", ddpm_schedule="log")
    with pytest.raises(IndexError):
        sampler.Sampler(None, a, scheduler.Scheduler(a), [None, None, None])._get_latent_initial(None)


@pytest.mark.parametrize("base", [2.0, 10.0])
def test_loss_weight_matches_reference(base):
    a = default_args(data_size=32, in_channel=3, out_channel=3, ddpm_num_steps=1000, select_degrade_pixel="indexing", ddpm_schedule="log")
    a.weight_dtype = torch.float32
    S = scheduler.Scheduler(a)
    assert S.update_ddpm_num_steps(1000) == int(Z["weight.Tp"][0])
    w = S.get_weight_timesteps(torch.from_numpy(Z["weight.idx"]), base)
    assert np.array_equal(w.numpy(), Z[f"weight.base{base}"])
