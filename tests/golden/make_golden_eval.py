"""Golden fixtures for the visual / evaluation side, produced by RUNNING THE REFERENCE:
  * `Sampler._save_image_grid` / `_save_multi_index_image_grid` (sampler.py:369-417: normalize01 / normalize01_global +
    torchvision make_grid) on seeded batches, incl. a grey batch, a constant image (0/0 -> NaN -> 0) and a ragged last row;
  * `Tester._compute_similarity`, `get_nearest_neighbor_idx`, `remove_duplicates_in_batches`,
    `remove_duplicates_across_batches` (tester.py:136-206) on a seeded data set with planted near-duplicates.
Run in the build container only:  python tests/golden/make_golden_eval.py  ->  tests/golden/eval.npz"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_shims import import_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def grid_inputs():
    g = torch.Generator().manual_seed(3)
    a = torch.randn(7, 3, 8, 8, generator=g)                 # ragged last row (nrow 3)
    b = torch.rand(4, 1, 6, 6, generator=g) * 2 - 1         # grey: make_grid replicates to 3 channels
    b[2] = 0.25                                              # constant image: normalize01 gives 0/0 -> 0
    c = torch.randn(2, 5, 3, 8, 8, generator=g)             # (batch, timesteps, C, H, W)
    return a, b, c


def eval_inputs():
    g = torch.Generator().manual_seed(5)
    data = torch.rand(40, 3, 8, 8, generator=g) * 2 - 1
    src = torch.rand(6, 3, 8, 8, generator=g)
    src[1] = (data[17] - data[17].min()) / (data[17].max() - data[17].min()) + 0.01 * torch.randn(3, 8, 8, generator=g)   # near data[17]
    src[4] = (data[3] - data[3].min()) / (data[3].max() - data[3].min())
    batch = torch.rand(8, 3, 8, 8, generator=g)
    batch[3] = batch[0] + 0.02 * torch.randn(3, 8, 8, generator=g)       # duplicates of earlier images
    batch[6] = batch[2] * 1.5
    prev = torch.stack([batch[5] + 0.01 * torch.randn(3, 8, 8, generator=g), torch.rand(3, 8, 8, generator=g)])
    return data, src, batch, prev


def main():
    ref_samp, ref_test = import_reference("sampler", "tester")
    out = {}
    a, b, c = grid_inputs()
    S = ref_samp.Sampler.__new__(ref_samp.Sampler)
    for nm, x in (("a", a), ("b", b)):
        for norm in ("global", "image"):
            out[f"grid/{nm}/{norm}"] = S._save_image_grid(x.clone(), normalization=norm).numpy()
    for norm in ("global", "image", None):
        for opt in (None, "skip_first"):
            grids = S._save_multi_index_image_grid(c.clone(), nrow=None, normalization=norm, option=opt)
            out[f"multigrid/{norm}/{opt}"] = torch.stack(grids).numpy()
    data, src, batch, prev = eval_inputs()
    T = ref_test.Tester.__new__(ref_test.Tester)
    T.cosine_similarity_th = 0.9
    T.args = SimpleNamespace(sample_num=16, data_size=8)
    T.dataset = [(data[i], 0) for i in range(data.shape[0])]
    out["eval/scores"] = T._compute_similarity(src, data, "cosine").numpy()          # (targets, sources), raw data
    out["eval/nn_idx"] = T.get_nearest_neighbor_idx(src).numpy()
    out["eval/dedup_in"] = T.remove_duplicates_in_batches(batch).numpy()
    out["eval/dedup_across"] = T.remove_duplicates_across_batches(batch, list(prev)).numpy()
    np.savez_compressed(os.path.join(OUT, "eval.npz"), **out)
    print({k: v.shape for k, v in out.items()})
    print("nn idx", out["eval/nn_idx"], "kept in batch", out["eval/dedup_in"].shape[0], "kept across", out["eval/dedup_across"].shape[0])


if __name__ == "__main__":
    main()
