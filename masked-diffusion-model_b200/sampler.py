"""Drop-in for the reference's `sampler.py` (class `Sampler`, /root/reference/code/sampler.py:
28-261): initial latent, iterative unmasking loop, `(sample_0, visual_list)` return.

Hot loop (sampler.py:137-258) on B200:
  * the data-independent draws of an iteration (shift noise of the NEXT iteration, the two
    masks of THIS iteration) are produced by the device mt19937 stream on a side CUDA stream
    while the denoiser runs on the main stream -- same word order as the reference
    (shift, degrade(t), degrade(t-1));
  * everything after the denoiser call -- x0_hat = (x_t+shift)+net-shift, both degradations,
    x_{t-1} = x_t + D(t-1) - D(t), and the next denoiser input x_{t-1} + shift' -- is the fused
    K5 update (csrc/degrade.cu), two launches per iteration, no separate elementwise ops.

The 11 per-step history tensors of the reference (sampler.py:116-126; 554 GB of host memory at
128x128 / 1000 steps / 256 images, quirk q20) are opt-in: `args.sample_history = True` fills
them; otherwise `visual_list` keeps its 11 slots with `None`.
Unsupported reference branches raise the same errors the reference raises (q11, q12)."""
from __future__ import annotations

import os

import numpy as np
import torch

from mdm_b200 import _lib
from mdm_b200._lib import check, lib, ptr, stream_ptr

HISTORY_NAMES = ["sample_t_list", "shift_list", "shifted_list", "mask_list", "shifted_result_list",
                 "sample_0_list", "degraded_mask_list", "degraded_mask_next_list", "degraded_t_list",
                 "difference_list", "degraded_next_t_list"]


def _strides_for(shift: torch.Tensor, B, C, H, W):
    """element strides (b, c, pixel) of a shift tensor broadcast to (B,C,H,W); pixel stride is
    1 for a full H*W plane, 0 for a broadcast one."""
    s = shift
    while s.dim() < 4:
        s = s[..., None]
    s = s.expand(B, C, H, W)
    sb, sc, sh, sw = s.stride()
    if sh == 0 and sw == 0:
        sp = 0
    elif sw == 1 and sh == W:
        sp = 1
    else:
        s = s.contiguous()
        sb, sc, sh, sw = s.stride()
        sp = 1
    return s, int(sb), int(sc), int(sp)


class Sampler:
    def __init__(self, dataset, args, Scheduler, dataset_hist=None):
        # the reference's base trainer calls this with 3 arguments (quirk q2): accepted here.
        self.dataset = dataset
        self.args = args
        self.Scheduler = Scheduler
        self.dataset_hist = dataset_hist
        self._side = None
        self._ws = {}

    # -- sampler.py:46-83 (CPU draws, exactly as the reference: they come first in the stream) --
    def _get_latent_initial(self, model):
        a = self.args
        if a.mean_area == "image-wise":
            ch = 1
        elif a.mean_area == "channel-wise":
            ch = 3
        shape = a.sample_latent_shape.lower()
        if shape == "data":
            hist_shape, edges, cum = self.dataset_hist[0], self.dataset_hist[1], self.dataset_hist[2]
            u = torch.rand(a.sample_num)
            idx = np.unravel_index(torch.searchsorted(cum, u).numpy(), hist_shape)
            cols = []
            for c in range(ch):
                v = torch.rand(a.sample_num)
                lo, hi = edges[c][idx[c]], edges[c][idx[c] + 1]
                cols.append(((hi - lo) * v + lo).unsqueeze(-1))
            sample_mean = torch.cat(cols, 1)
        elif shape == "zero":
            sample_mean = torch.zeros(a.sample_num, ch)
        elif shape == "normal":
            sample_mean = torch.randn(a.sample_num, ch)
        elif shape == "uniform":
            sample_mean = torch.FloatTensor(a.sample_num, ch).uniform_(-1, 1)
        elif shape == "grid":
            raise IndexError("too many indices for tensor of dimension 1")   # quirk q12
        sample = sample_mean[:, :, None, None]
        return sample.expand(a.sample_num, a.out_channel, a.data_size, a.data_size)

    def sample(self, model, timesteps_used_epoch, interpolation_shift=None):
        return self._sample_mean_shift_momentum(model, timesteps_used_epoch)

    # -- helpers ---------------------------------------------------------------------------------
    def _buf(self, name, shape, dtype, device):
        key = (name, tuple(shape), dtype, str(device))
        t = self._ws.get(key)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=device)
            self._ws[key] = t
        return t

    def _draw_shift(self, S, time, like):
        """shift tensor for timestep tensor `time` (consumes the stream like scheduler.py:612)."""
        return S.get_schedule_shift_time(time, like)

    # -- sampler.py:109-261 --------------------------------------------------------------------------
    def _sample_mean_shift_momentum(self, model, timesteps_used_epoch):
        a, S = self.args, self.Scheduler
        dev = torch.device(model.device)
        if dev.type != "cuda":
            raise RuntimeError("Sampler needs the denoiser on a CUDA device (no CPU fallback)")
        if a.momentum_adaptive in ("momentum", "boosting"):
            raise UnboundLocalError("cannot access local variable 'momentum' where it is not associated with a value")  # q11
        if a.momentum_adaptive not in ("base_momentum", "base_sampling"):
            raise ValueError(a.momentum_adaptive)
        dep = a.sampling_mask_dependency
        if dep not in ("independent", "dependent_prev", "dependent_t"):
            raise ValueError(dep)

        N, C, H = a.sample_num, a.out_channel, a.data_size
        hw = H * H
        T = len(timesteps_used_epoch)
        latent = self._get_latent_initial(model)              # CPU generator draws first
        S.adopt_torch_rng(dev)                                # then the stream moves to the device
        rng = S.rng
        x_t = latent.to(dev, dtype=torch.float32).contiguous()
        history = bool(getattr(a, "sample_history", False))
        hist = None
        if history:
            # the eleven history tensors of sampler.py:116-126: recorded on the device (no per-step synchronising
            # device-to-host copies) when they fit the budget, handed back as CPU tensors like the reference's
            hist_bytes = len(HISTORY_NAMES) * (T + 1) * N * C * H * H * 4
            hist_dev = dev if hist_bytes <= int(os.environ.get("MDM_HISTORY_DEVICE_BYTES", str(8 << 30))) else "cpu"
            hist = [torch.zeros(T + 1, N, C, H, H, device=hist_dev) for _ in HISTORY_NAMES]

        mask_ch = S._mask_channels()
        mode, const, area = S._fill_mode(a.mean_option, a.mean_area)
        momentum = 1 if a.momentum_adaptive == "base_momentum" else 0
        ws = self._buf("ws", (max(1, 2 * lib().mdm_degrade_ws_floats(N, C, hw)),), torch.float32, dev)
        mb_t = self._buf("mb_t", (N, mask_ch * hw), torch.uint8, dev)
        mb_n = [self._buf("mb_n0", (N, mask_ch * hw), torch.uint8, dev),
                self._buf("mb_n1", (N, mask_ch * hw), torch.uint8, dev)]
        mb_n[0].fill_(0)
        mb_n[1].fill_(0)
        zeros_like = torch.zeros(1, 1, 1, 1, device=dev).expand(N, C, H, H)

        main = torch.cuda.current_stream(dev)
        if self._side is None:
            self._side = torch.cuda.Stream(device=dev)
        side = self._side
        side.wait_stream(main)

        def time_tensor(t):
            return torch.full((N,), float(t), dtype=torch.float32, device=dev)

        # shift of the first iteration (stream order: shift, degrade(t), degrade(t-1), shift', ...)
        with torch.cuda.stream(side):
            time = time_tensor(timesteps_used_epoch[T - 1])
            shift = self._draw_shift(S, time, zeros_like)
            shift_e, sb, sc, sp = _strides_for(shift.float(), N, C, H, H)
        main.wait_stream(side)
        x_in = torch.empty_like(x_t)
        check(lib().mdm_add_shift(ptr(x_t), ptr(shift_e), sb, sc, sp, ptr(x_in), N, C, hw, stream_ptr(dev)))
        sample_0 = torch.empty_like(x_t)
        x_next = torch.empty_like(x_t)
        x_in_next = torch.empty_like(x_t)
        difference_prev = None

        # the mask / noise generator is a serial one-CTA kernel that runs on the side stream for about half of every
        # denoising step.  With STATIC work lists (MDM_IGEMM_DYNAMIC=0) every persistent GEMM launched meanwhile waits
        # for the CTA that shares its SM, so the generator gets an SM of its own (44.5 -> 41.2 ms per step at
        # 256x3x128x128); with the dynamic work distribution (the default) that CTA just takes fewer items and no SM
        # has to stay idle.
        from mdm_b200 import denoiser_ops as _dops
        static_lists = os.environ.get("MDM_IGEMM_DYNAMIC", "1") == "0"
        reserved_before = _dops.reserve_sms(int(os.environ.get("MDM_SAMPLER_RESERVE_SMS", "1" if static_lists else "0")))
        # the loop's mask / noise draws run UNDER the denoiser: fewer, longer pieces per draw (every CTA of a parallel
        # draw pays the same jump) -- 34.77 -> 34.38 ms per step at 256x3x128x128
        stride_before = lib().mdm_rng_set_par_stride(int(os.environ.get("MDM_SAMPLER_RNG_STRIDE", "8")))
        try:
            with torch.no_grad():
                for i in range(T - 1, -1, -1):
                    t = timesteps_used_epoch[i]
                    time = time_tensor(t)
                    next_t = time - 1 if i > 0 else time
                    # ---- side stream: masks of this iteration, shift of the next one ---------------
                    side.wait_stream(main)
                    with torch.cuda.stream(side):
                        n_t = S.get_black_area_num_pixels_time(time)
                        n_next = S.get_black_area_num_pixels_time(next_t)
                        cur_n = mb_n[i & 1]
                        prev_n = mb_n[(i + 1) & 1]
                        if dep == "independent":
                            S.make_mask_bytes(n_t, dev, out=mb_t)
                            S.make_mask_bytes(n_next, dev, out=cur_n)
                            m_t_bytes = mb_t
                        elif dep == "dependent_prev":
                            S.make_mask_bytes(n_next, dev, out=cur_n)
                            m_t_bytes = prev_n                      # mask drawn for t in the previous iteration
                        else:  # dependent_t: one uniform field, two thresholds
                            if a.select_degrade_pixel != "thresholding":
                                raise UnboundLocalError("masks_t")
                            rng.threshold_mask(n_t.double(), N, mask_ch * hw, ratio2=n_next.double(), out=mb_t, out2=cur_n)
                            m_t_bytes = mb_t
                        if i > 0:
                            time_n = time_tensor(timesteps_used_epoch[i - 1])
                            shift_n = self._draw_shift(S, time_n, zeros_like)
                            shift_ne, nb, nc, np_ = _strides_for(shift_n.float(), N, C, H, H)
                        else:
                            shift_ne, nb, nc, np_ = None, 0, 0, 0
                    # ---- main stream: denoiser ------------------------------------------------------
                    net = model(x_in, time).sample
                    if net.dtype != torch.float32 or not net.is_contiguous():
                        net = net.float().contiguous()
                    main.wait_stream(side)
                    # ---- fused update (K5) ----------------------------------------------------------
                    last = (i == 0)
                    update = 0 if (last and momentum) else 1       # sampler.py:204-216
                    if last and not momentum:
                        update = 0                                  # base_sampling breaks before the update
                    check(lib().mdm_sampler_step(
                        ptr(x_t), ptr(net), ptr(shift_e), sb, sc, sp, ptr(m_t_bytes), ptr(cur_n), mask_ch,
                        mode, const, area, momentum, update, ptr(shift_ne), nb, nc, np_,
                        ptr(x_next), ptr(x_in_next) if not last else None,
                        ptr(sample_0) if (last or history) else None,      # x0_hat is only read after the last iteration
                        ptr(ws), N, C, hw,
                        stream_ptr(dev)))
                    if history:
                        difference_prev = self._record(hist, T - i, x_t, shift_e, x_in, net, sample_0, m_t_bytes, cur_n,
                                                       mask_ch, mode, const, area, dep, last, momentum, difference_prev)
                    x_t, x_next = x_next, x_t
                    x_in, x_in_next = x_in_next, x_in
                    shift_e, sb, sc, sp = shift_ne, nb, nc, np_
        finally:
            _dops.reserve_sms(reserved_before)
            lib().mdm_rng_set_par_stride(stride_before)
        S.release_rng_to_torch()
        visual = [h.cpu() for h in hist] if history else [None] * len(HISTORY_NAMES)
        return sample_0, visual

    # -- image grids on the device (sampler.py:369-417) ------------------------------------------------------
    @staticmethod
    def _grid(sample: torch.Tensor, nrow: int, normalization) -> torch.Tensor:
        """`make_grid(normalize01[_global](sample), nrow)` in one pass on the device (csrc/grid.cu)"""
        if not sample.is_cuda:
            raise RuntimeError("image grids are built on the device (no CPU fallback)")
        x = sample.float().contiguous()
        B, C, H, W = x.shape
        mode = {None: 0, "image": 1, "global": 2}[normalization]
        xm = min(nrow, B)
        ym = (B + xm - 1) // xm
        pad = 2
        out = torch.empty(3 if C == 1 else C, (H + pad) * ym + pad, (W + pad) * xm + pad, device=x.device)
        ws = torch.empty(2 * B, device=x.device)
        check(lib().mdm_image_grid(ptr(x), B, C, H, W, nrow, pad, 0.0, mode, ptr(ws), ptr(out), stream_ptr(x.device)))
        return out

    def _save_image_grid(self, sample: torch.Tensor, normalization='global', dir_save=None, file_sample=None):
        """sampler.py:369-388: square grid of a batch; written as an image file when a path is given"""
        batch_size = sample.shape[0]
        nrow = int(np.ceil(np.sqrt(batch_size)))
        if nrow == 0:
            return None                                  # the reference's ZeroDivisionError branch
        grid_sample = self._grid(sample, nrow, normalization if normalization in ("global", "image") else None)
        if dir_save is not None and file_sample is not None:
            from torchvision.utils import save_image      # host-side PNG encoding only
            save_image(grid_sample, os.path.join(dir_save, file_sample))
        return grid_sample

    def _save_multi_index_image_grid(self, sample: torch.Tensor, nrow=None, normalization='global', option=None):
        """sampler.py:391-417: one grid per batch element over its timesteps; sample: (batch, timesteps, C, H, W)"""
        num_timesteps = sample.shape[1]
        if nrow is None:
            nrow = int(np.ceil(np.sqrt(num_timesteps)))
        grids = []
        for i in range(sample.shape[0]):
            s_i = sample[i][1:] if option == 'skip_first' else sample[i]
            grids.append(self._grid(s_i.to("cuda") if not s_i.is_cuda else s_i, nrow,
                                    normalization if normalization in ("global", "image") else None))
        return grids

    # -- opt-in history (visualisation; not on the timed path) ----------------------------------
    def _record(self, hist, k, x_t, shift, x_in, net, s0, mb_t, mb_n, mask_ch, mode, const, area, dep, last,
                momentum, difference_prev):
        S = self.Scheduler
        N, C, H = x_t.shape[0], x_t.shape[1], x_t.shape[2]
        full = lambda t: t.expand(N, C, H, H) if t is not None else torch.zeros(N, C, H, H, device=x_t.device)
        m_t = mb_t.view(N, mask_ch, H, H).float()
        m_n = mb_n.view(N, mask_ch, H, H).float()
        d_t = S.degrade_with_mask(s0, m_t, self.args.mean_option, self.args.mean_area)
        d_n = S.degrade_with_mask(s0, m_n, self.args.mean_option, self.args.mean_area)
        hist[0][k] = x_t
        hist[1][k] = full(shift)
        hist[2][k] = x_in
        hist[3][k] = net
        hist[4][k] = (x_in + net)
        hist[5][k] = s0
        if dep == "dependent_prev":
            hist[6][k] = full(m_n)
        else:
            hist[6][k] = full(m_t)
            hist[7][k] = full(m_n)
        if last and not momentum:
            return difference_prev                      # base_sampling breaks before recording
        hist[8][k] = d_t
        hist[10][k] = d_n
        difference = (d_n - d_t) if not (last and momentum) else difference_prev   # quirk q21
        if difference is not None:
            hist[9][k] = difference
        return difference
