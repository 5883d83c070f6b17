"""CPU: the WHOLE denoiser oracle (oracle/unet_ref.py) against the reference's own in-repo network of the same
topology -- `models/unet/unet6.py:365-506` as `UNet(C, 128, C, (1,1,2,2,4,4), 2, (F,F,F,F,T,F))`
(`models_Unet.py:153-159`) -- run by tests/golden/make_golden_unet6.py with seeded weights.  The four places where
that network differs from the diffusers defaults are diffusers config fields the oracle honours
(`unet6_compat()`: norm_eps 1e-6, [sin, cos] / half - 1 embedding, one 1/sqrt(C) attention head, right / bottom
stride-2 padding); with them set, the oracle must reproduce the reference's output to fp32 rounding.
Tolerance: relative L2 <= 1e-5 (fp32 CPU on both sides; summation orders differ between the 1x1-conv and Linear
formulations of the attention projections)."""
import numpy as np
import pytest
import torch

from oracle.unet_ref import (UNet2DModelRef, seeded_state_dict, timestep_embedding, to_unet6_state_dict, unet6_compat,
                             unet6_key_map, unet_config)
from tests.golden.make_golden_unet6 import CASES


@pytest.mark.parametrize("name", list(CASES))
def test_whole_network_matches_reference_unet6(golden, name):
    g = golden("unet6_net")
    c = CASES[name]
    cfg = unet_config(c["C"], c["S"])
    cfg.update(unet6_compat(cfg["block_out_channels"][-1]))
    net = UNet2DModelRef(**cfg).eval()
    sd = seeded_state_dict(net, c["seed"])
    assert sum(v.numel() for v in sd.values()) == int(g[f"{name}/n_params"])
    flat = torch.cat([v.flatten()[:: max(1, v.numel() // 64)][:64].double() for v in sd.values()])
    fp = np.array([float(flat.sum()), float(flat.abs().sum()), float(flat[12345 % flat.numel()])])
    np.testing.assert_allclose(fp, g[f"{name}/w_fingerprint"], rtol=1e-12)       # same seeded weights as the generator's
    net.load_state_dict(sd)
    with torch.no_grad():
        y = net(torch.from_numpy(g[f"{name}/x"]), torch.from_numpy(g[f"{name}/t"])).sample
    want = torch.from_numpy(g[f"{name}/y"])
    rel = ((y - want).norm() / want.norm()).item()
    print(f"{name}: oracle vs reference unet6, rel L2 = {rel:.2e}")
    assert rel <= 1e-5, rel


def test_each_compat_field_matters(golden):
    """dropping any one of the four compat fields moves the output far outside the tolerance: the pin is sensitive to
    eps, embedding layout, head split and padding side individually (the defaults are what diffusers documents)"""
    g = golden("unet6_net")
    name = "c3_s32"
    c = CASES[name]
    want = torch.from_numpy(g[f"{name}/y"])
    base = unet_config(c["C"], c["S"])
    compat = unet6_compat(base["block_out_channels"][-1])
    for drop in (("norm_eps",), ("flip_sin_to_cos", "freq_shift"), ("attention_head_dim",), ("downsample_padding",)):
        cfg = dict(base)
        cfg.update({k: v for k, v in compat.items() if k not in drop})
        net = UNet2DModelRef(**cfg).eval()
        net.load_state_dict(seeded_state_dict(net, c["seed"]))
        with torch.no_grad():
            y = net(torch.from_numpy(g[f"{name}/x"]), torch.from_numpy(g[f"{name}/t"])).sample
        rel = ((y - want).norm() / want.norm()).item()
        assert rel > (1e-5 if drop == ("norm_eps",) else 1e-3), (drop, rel)


def test_key_map_covers_every_parameter():
    cfg = unet_config(3, 32)
    net = UNet2DModelRef(**cfg)
    sd = net.state_dict()
    mapped = to_unet6_state_dict(sd, cfg)
    assert sum(v.numel() for v in mapped.values()) == sum(v.numel() for v in sd.values()) == 113_673_219
    km = unet6_key_map(cfg)
    assert km["down_blocks.4.attentions.1.qkv"] == "downsamples.level_4.1.1.project_in"
    assert km["up_blocks.1.resnets.2.conv_shortcut"] == "upsamples.level_4.2.0.skip"
    assert km["up_blocks.4.upsamplers.0.conv"] == "upsamples.level_1.3.1"


def test_timestep_embedding_variants():
    t = torch.tensor([0.0, 1.0, 999.0])
    d = timestep_embedding(t, 128)                                   # diffusers default: [cos, sin], / half
    u = timestep_embedding(t, 128, flip_sin_to_cos=False, freq_shift=1)
    assert torch.equal(d[0, :64], torch.ones(64)) and torch.equal(u[0, 64:], torch.ones(64))
    assert abs(float(d[1, 63 + 64]) - np.sin(10000.0 ** (-63 / 64))) < 1e-6
    assert abs(float(u[1, 63]) - np.sin(10000.0 ** (-63 / 63))) < 1e-6
