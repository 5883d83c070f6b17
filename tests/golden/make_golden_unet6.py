"""Golden fixture that pins the WHOLE denoiser oracle (oracle/unet_ref.py) to reference-held code.

Runs the reference's in-repo `models/unet/unet6.py:365-506` -- `UNet(C, 128, C, (1,1,2,2,4,4), 2, (F,F,F,F,T,F))`, the
configuration of `models_Unet.py:153-159`, same topology as the diffusers network of `utils/model.py` -- on CPU with
SEEDED weights (`oracle.unet_ref.seeded_state_dict`, re-keyed by `to_unet6_state_dict`; the 447 MB of weights are
regenerated from the seed by the test and never stored), and stores input, timesteps and output in
tests/golden/unet6_net.npz together with a checksum of the weights.

Run in the build container only (the reference tree does not exist on the GPU box):
    python tests/golden/make_golden_unet6.py"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("MDM_REFERENCE_CODE", "/root/reference/code")
OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {"c3_s32": dict(C=3, S=32, B=2, seed=7, t=[3.0, 977.0]), "c1_s32": dict(C=1, S=32, B=2, seed=8, t=[250.0, 1.0]),
         "c3_s64": dict(C=3, S=64, B=1, seed=9, t=[500.0])}


def main():
    from tests.golden.make_golden_blocks import load
    from oracle.ref_shims import install_stubs
    from oracle.unet_ref import UNet2DModelRef, seeded_state_dict, to_unet6_state_dict, unet6_compat, unet_config
    install_stubs()
    u6 = load(os.path.join(REF, "models", "unet", "unet6.py"), "ref_unet6")
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    out = {}
    for name, c in CASES.items():
        cfg = unet_config(c["C"], c["S"])
        cfg.update(unet6_compat(cfg["block_out_channels"][-1]))
        oracle = UNet2DModelRef(**cfg)
        sd = seeded_state_dict(oracle, c["seed"])
        net = u6.UNet(c["C"], 128, c["C"], (1, 1, 2, 2, 4, 4), 2, (False, False, False, False, True, False)).eval()
        missing = net.load_state_dict(to_unet6_state_dict(sd, cfg), strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        assert sum(p.numel() for p in net.parameters()) == sum(v.numel() for v in sd.values())
        g = torch.Generator().manual_seed(c["seed"] + 100)
        x = torch.rand(c["B"], c["C"], c["S"], c["S"], generator=g) * 2 - 1
        t = torch.tensor(c["t"])
        with torch.no_grad():
            y = net(x, t)
        out[f"{name}/x"], out[f"{name}/t"], out[f"{name}/y"] = x.numpy(), t.numpy(), y.numpy()
        # weights fingerprint: detects a change of the seeded generator between the machine that made the fixture and
        # the one that checks it
        flat = torch.cat([v.flatten()[:: max(1, v.numel() // 64)][:64].double() for v in sd.values()])
        out[f"{name}/w_fingerprint"] = np.array([float(flat.sum()), float(flat.abs().sum()), float(flat[12345 % flat.numel()])])
        out[f"{name}/n_params"] = np.int64(sum(v.numel() for v in sd.values()))
        print(name, "params", int(out[f"{name}/n_params"]), "out rms", float(y.pow(2).mean().sqrt()))
    np.savez_compressed(os.path.join(OUT, "unet6_net.npz"), **out)


if __name__ == "__main__":
    main()
