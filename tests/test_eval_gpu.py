"""GPU: the visual / evaluation side against fixtures produced by the reference's own code
(tests/golden/make_golden_eval.py -> eval.npz): device-built image grids vs `Sampler._save_image_grid` /
`_save_multi_index_image_grid` (sampler.py:369-417), and the tensor-core cosine GEMM behind `Tester._compute_similarity`
/ `get_nearest_neighbor_idx` / `remove_duplicates_*` (tester.py:136-206).
Tolerances: grids are the reference's fp32 arithmetic op for op -> 1e-6; cosine scores from the three-piece bf16
split GEMM -> 2e-6 absolute; indices and kept sets exact."""
import numpy as np
import pytest
import torch

import sampler
from tests.golden.make_golden_eval import eval_inputs, grid_inputs

pytestmark = pytest.mark.gpu


def test_image_grids_match_reference(golden):
    g = golden("eval")
    a, b, c = grid_inputs()
    S = sampler.Sampler(None, None, None, None)
    for nm, x in (("a", a), ("b", b)):
        for norm in ("global", "image"):
            got = S._save_image_grid(x.cuda(), normalization=norm)
            want = g[f"grid/{nm}/{norm}"]
            assert tuple(got.shape) == want.shape
            np.testing.assert_allclose(got.cpu().numpy(), want, atol=1e-6, rtol=0, err_msg=f"{nm}/{norm}")
    for norm in ("global", "image", None):
        for opt in (None, "skip_first"):
            grids = S._save_multi_index_image_grid(c.cuda(), nrow=None, normalization=norm, option=opt)
            want = g[f"multigrid/{norm}/{opt}"]
            assert len(grids) == want.shape[0]
            np.testing.assert_allclose(torch.stack(grids).cpu().numpy(), want, atol=1e-6, rtol=0, err_msg=f"{norm}/{opt}")


def test_image_grid_file_output(tmp_path):
    a, _, _ = grid_inputs()
    S = sampler.Sampler(None, None, None, None)
    grid = S._save_image_grid(a.cuda(), normalization="image", dir_save=str(tmp_path), file_sample="grid.png")
    assert (tmp_path / "grid.png").stat().st_size > 100 and grid.shape == (3, 32, 32)


def test_cosine_gemm_and_neighbour_search_match_reference(golden):
    import tester
    from types import SimpleNamespace
    g = golden("eval")
    data, src, batch, prev = eval_inputs()
    T = tester.Tester(SimpleNamespace(sample_num=16, data_size=8), [(data[i], 0) for i in range(data.shape[0])])
    score = T._compute_similarity(src.cuda(), data.cuda(), "cosine")
    assert score.shape == (40, 6)
    np.testing.assert_allclose(score.cpu().numpy(), g["eval/scores"], atol=2e-6, rtol=0)
    assert np.array_equal(T.get_nearest_neighbor_idx(src.cuda()).cpu().numpy(), g["eval/nn_idx"])
    assert np.array_equal(T.get_nearest_neighbor_idx(src.cuda(), batch=16).cpu().numpy(), g["eval/nn_idx"])     # chunked data set
    kept = T.remove_duplicates_in_batches(batch.cuda())
    assert np.array_equal(kept.cpu().numpy(), g["eval/dedup_in"])
    across = T.remove_duplicates_across_batches(batch.cuda(), list(prev))
    assert np.array_equal(across.cpu().numpy(), g["eval/dedup_across"])


def test_cosine_gemm_at_image_scale():
    """3 x 32 x 32 images (K = 3072 per piece, 18432 after the six-way concatenation), 300 targets x 100 sources: the
    GEMM scores stay within 2e-6 of torch's fp32 cosine similarity"""
    import tester
    g = torch.Generator(device="cuda").manual_seed(1)
    src = torch.rand(100, 3, 32, 32, device="cuda", generator=g)
    tgt = torch.rand(300, 3, 32, 32, device="cuda", generator=g)
    got = tester.cosine_scores(src, tgt)
    want = torch.nn.functional.cosine_similarity(src.flatten(1)[None].double(), tgt.flatten(1)[:, None].double(), dim=2)
    assert (got.double() - want).abs().max().item() <= 2e-6
