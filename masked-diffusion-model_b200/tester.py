"""Drop-in for the evaluation helpers of the reference's `tester.py` (class `Tester`, /root/reference/code/tester.py:
136-230): cosine similarity between generated images and a data set, nearest-neighbour lookup, duplicate removal.

The reference computes `cosine_similarity(source[None], target[:, None], dim=2)` per data-loader batch -- a
(targets x sources x pixels) broadcast -- and one Python-level `cosine_similarity` call per image pair in the
de-duplication loops.  Here the score matrix is ONE tensor-core GEMM (the implicit-GEMM kernel of the denoiser,
csrc/igemm.cu, used as a linear layer): rows are L2-normalised in fp32 and split into three bf16 pieces
(x = hi + mid + lo, 24 mantissa bits), and the K-concatenated product [hi | mid | hi | lo | mid | hi] x [hi | hi | mid | hi | mid | lo]
recovers the fp32 dot product to ~1e-6; arg-max candidates are re-scored exactly in fp32.  The greedy keep / drop
decisions of the de-duplication then read the matrix on the host (a few hundred entries).

Out of scope: plotting, image files, the sampling driver of the reference's `Tester.train`."""
from __future__ import annotations

import torch

from mdm_b200 import denoiser_ops as ops


def normalize01(data: torch.Tensor) -> torch.Tensor:
    """utils/datautils.py:211-222"""
    b = data.shape[0]
    mx = torch.amax(data, dim=(1, 2, 3)).reshape(b, 1, 1, 1)
    mn = torch.amin(data, dim=(1, 2, 3)).reshape(b, 1, 1, 1)
    return torch.nan_to_num((data - mn) / (mx - mn), nan=0.0)


def _split3(x: torch.Tensor):
    hi = x.to(torch.bfloat16)
    r = x - hi.float()
    mid = r.to(torch.bfloat16)
    lo = (r - mid.float()).to(torch.bfloat16)
    return hi, mid, lo


def cosine_scores(source: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """score[t, s] = cos(target[t], source[s]) over the flattened images -- `Tester._compute_similarity(source, target)`
    (tester.py:142-147) as one tcgen05 GEMM.  source: (S, ...), target: (T, ...) CUDA tensors; returns (T, S) fp32."""
    if not (source.is_cuda and target.is_cuda):
        raise RuntimeError("cosine_scores: CUDA tensors only (no CPU fallback)")
    vs = source.flatten(1).float()
    vt = target.flatten(1).float()
    S, D = vs.shape
    T = vt.shape[0]
    eps = 1e-8                                                 # F.cosine_similarity: x / max(||x||, eps)
    vs = vs / vs.norm(dim=1, keepdim=True).clamp_min(eps)
    vt = vt / vt.norm(dim=1, keepdim=True).clamp_min(eps)
    Dp = (D + 63) // 64 * 64
    Sp = (S + 31) // 32 * 32

    def pad(x, rows):
        out = torch.zeros(rows, Dp, device=x.device)
        out[:x.shape[0], :D] = x
        return out
    th, tm, tl = _split3(pad(vt, T))
    sh, sm, sl = _split3(pad(vs, Sp))
    # sum of the six products of magnitude >= 2^-16: hi*hi + mid*hi + hi*mid + lo*hi + mid*mid + hi*lo
    A = torch.cat([th, tm, th, tl, tm, th], dim=1).contiguous()            # [T, 6 Dp]
    Bm = torch.cat([sh, sh, sm, sh, sm, sl], dim=1).contiguous()           # [Sp, 6 Dp]
    out = torch.empty(T, Sp, dtype=torch.float32, device=vs.device)
    ops.conv_fprop(A, Bm.view(Sp, 1, 6 * Dp), None, T, 1, 1, 1, 1, y_f32=out, cout=Sp)
    return out[:, :S]


class Tester:
    """the evaluation helpers of tester.py:136-230 (same method names); construct with `(args, dataset)`"""

    def __init__(self, args=None, dataset=None):
        self.args, self.dataset = args, dataset
        self.cosine_similarity_th = 0.9                        # tester.py:53

    def cosine_similarity(self, image1, image2):               # tester.py:136-139
        return torch.nn.functional.cosine_similarity(image1.flatten().float(), image2.flatten().float(), dim=0)

    def _compute_similarity(self, source: torch.Tensor, target: torch.Tensor, metric: str = 'cosine'):
        if metric.lower() != 'cosine':
            raise UnboundLocalError("score")                   # the reference only assigns `score` for 'cosine'
        return cosine_scores(source, target)

    def get_nearest_neighbor_idx(self, source: torch.Tensor, batch: int = 8192):
        """tester.py:189-206: index of the most similar (normalize01-ed) data-set image for every generated image.
        The data set is streamed in chunks; per chunk one GEMM, the running arg-max is re-scored exactly in fp32."""
        dev = source.device
        S = source.shape[0]
        best_val = torch.full((S,), -float("inf"), device=dev)
        best_idx = torch.zeros(S, dtype=torch.long, device=dev)
        src = source.flatten(1).float()
        n = len(self.dataset)
        for lo in range(0, n, batch):
            items = [self.dataset[i] for i in range(lo, min(lo + batch, n))]
            data = torch.stack([it[0] if isinstance(it, (tuple, list)) else it for it in items]).to(dev)
            data = normalize01(data.float())
            score = cosine_scores(source, data)                               # (chunk, S)
            k = min(4, score.shape[0])
            cand = score.topk(k, dim=0).indices                               # (k, S) near-best candidates of the chunk
            d = data.flatten(1)[cand]                                         # (k, S, D)
            exact = torch.nn.functional.cosine_similarity(src[None], d, dim=2)     # fp32, like the reference's entries
            val, j = exact.max(dim=0)
            idx = cand.gather(0, j[None])[0] + lo
            better = val > best_val                                           # first maximum wins, like torch.max over dim 0
            best_val = torch.where(better, val, best_val)
            best_idx = torch.where(better, idx, best_idx)
        return best_idx

    def remove_duplicates_in_batches(self, current_batch):
        """tester.py:150-162: greedy, keeps an image unless it is >= th similar to an image kept before it"""
        score = cosine_scores(current_batch, current_batch).cpu()
        keep = [0]
        for i in range(1, current_batch.shape[0]):
            if not any(float(score[j, i]) >= self.cosine_similarity_th for j in keep):
                keep.append(i)
        return current_batch[keep]

    def remove_duplicates_across_batches(self, unique_in_batch, previous_images):
        """tester.py:165-186: drops the images that are > th similar to any earlier image"""
        if len(previous_images) == 0:
            return unique_in_batch
        prev = previous_images if torch.is_tensor(previous_images) else torch.stack(list(previous_images))
        score = cosine_scores(unique_in_batch, prev.to(unique_in_batch.device))          # (prev, new)
        keep = ~(score > self.cosine_similarity_th).any(dim=0)
        if not bool(keep.any()):
            return torch.empty(0, *unique_in_batch.shape[1:], device=unique_in_batch.device)
        return unique_in_batch[keep]
