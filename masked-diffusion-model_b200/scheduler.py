"""Drop-in for the reference's `scheduler.py` (class `Scheduler`, /root/reference/code/
scheduler.py:13-794) on B200.

Same constructor, method names, argument meaning and return arity; the device work goes through
libmdm_sm100.so (include/mdm.h):

* masks come from a device-resident copy of torch's CPU mt19937 stream (csrc/rng.cu), so they
  are bit-identical to the reference's CPU-generated masks under the same seed;
* fill + composite is the fused K1 kernel pair (csrc/degrade.cu).

Differences that are deliberate: the per-call `torch.tensor(list,
device=...)` rebuild (q14) is replaced by a cached device table; unsupported reference branches
(`sigmoid` schedule q5, `indexing`+`linear` q4) raise the same exception types early.  There is
no CPU path: images must be CUDA tensors.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from mdm_b200 import _lib
from mdm_b200._lib import check, lib, ptr, require_cuda, stream_ptr
from mdm_b200.rng import DeviceMT19937


class Scheduler:
    def __init__(self, args):
        self.args = args
        self.height = args.data_size
        self.width = args.data_size
        self.image_size = self.height * self.width

        self.updated_ddpm_num_steps = None
        self.ratio_list = None
        self.black_area_pixels = None
        self.reverse_ratio = None

        self.rng: DeviceMT19937 | None = None     # device shadow of torch's CPU generator
        self._tables = {}                         # device -> (ratio f64, counts i64 | None)
        self._ws = {}                             # workspace cache

    # ------------------------------------------------------------------------------------------
    # schedule (host, one-off)  -- scheduler.py:27-65, 103-170
    # ------------------------------------------------------------------------------------------
    def update_ddpm_num_steps(self, max_time):
        kind = self.args.ddpm_schedule
        steps = self.args.ddpm_num_steps
        if kind == "linear":
            sched = self.get_extract_linear_random_sublist(steps)
        elif kind == "log":
            sched = self.get_extract_log_random_sublist(list(range(1, self.image_size + 1)), steps)
        elif kind == "exponential":
            sched = self.get_extract_exponential_random_sublist(steps)
        elif kind == "sigmoid":
            # the reference reaches torch.flip(ndarray) and raises TypeError (quirk q5)
            raise TypeError("ddpm_schedule='sigmoid' is broken in the reference (torch.flip on ndarray)")
        else:
            raise ValueError("Invalid mask ratio scheduler")

        if kind == "log":
            sched[-1] = self.image_size
            self.ratio_list = torch.tensor(sched / self.image_size)
        else:
            self.ratio_list = sched
        self.reverse_ratio = torch.flip(self.ratio_list, dims=(0,))
        self.black_area_pixels = sched
        self.updated_ddpm_num_steps = len(sched)
        self._tables.clear()
        return self.updated_ddpm_num_steps

    def get_black_area_num_pixels_all(self):
        return self.black_area_pixels

    def get_updated_ddpm_num_steps(self):
        return self.updated_ddpm_num_steps

    def get_ratio_list(self):
        return self.ratio_list

    def get_reverse_ratio_list(self):
        return self.reverse_ratio

    def get_extract_linear_random_sublist(self, ddpm_num_steps):
        return torch.tensor(np.linspace(1e-3, 1, ddpm_num_steps))

    def get_extract_log_random_sublist(self, time_list, ddpm_num_steps):
        if ddpm_num_steps > len(time_list):
            raise ValueError("Desired to remove number of pixels is greater than the size of input image.")
        v = np.log(np.linspace(1, self.image_size, ddpm_num_steps))
        v = v - min(v) + 1
        v = v * (self.image_size / max(v))
        return np.array(sorted(set(np.asarray(v, dtype=int))))

    def get_extract_exponential_random_sublist(self, ddpm_num_steps):
        e = self.args.ddpm_schedule_base ** np.linspace(0, 1, ddpm_num_steps)
        return torch.tensor(e / e[-1])

    def get_timesteps_epoch(self, epoch, epoch_length):
        scale = self.args.scheduler_num_scale_timesteps
        total = self.updated_ddpm_num_steps
        section = math.ceil((epoch + 1) / (epoch_length / scale))
        try:
            stride = np.power(2, scale - section)
            used = [i for i in range(1, total + 1) if i % stride == 0]
        except ValueError:            # negative integer power: every timestep
            used = list(range(1, total + 1))
        used[-1] = total
        return used

    # ------------------------------------------------------------------------------------------
    # device tables / RNG plumbing
    # ------------------------------------------------------------------------------------------
    def _table(self, device):
        key = str(device)
        if key not in self._tables:
            ratio = torch.as_tensor(self.ratio_list, dtype=torch.float64).to(device)
            counts = None
            if isinstance(self.black_area_pixels, np.ndarray):
                counts = torch.as_tensor(self.black_area_pixels, dtype=torch.int64).to(device)
            self._tables[key] = (ratio, counts)
        return self._tables[key]

    @staticmethod
    def _norm_device(device):
        d = torch.device(device)
        if d.type == "cuda" and d.index is None:
            d = torch.device("cuda", torch.cuda.current_device())
        return d

    def adopt_torch_rng(self, device, generator=None):
        """Copy torch's CPU generator state to the device; masks continue from there."""
        device = self._norm_device(device)
        if self.rng is None or self.rng.device != device:
            self.rng = DeviceMT19937(device)
        self.rng.adopt_torch(generator)
        return self.rng

    def release_rng_to_torch(self, generator=None):
        """Write the advanced state back so later CPU draws stay aligned with the reference."""
        if self.rng is not None:
            self.rng.release_to_torch(generator)

    def _rng_for(self, device):
        if self.rng is None or self.rng.device != self._norm_device(device):
            self.adopt_torch_rng(device)
        return self.rng

    def _buf(self, name, shape, dtype, device):
        key = (name, tuple(shape), dtype, str(device))
        t = self._ws.get(key)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=device)
            self._ws[key] = t
        return t

    # ------------------------------------------------------------------------------------------
    # scheduler.py:88-100
    # ------------------------------------------------------------------------------------------
    def get_black_area_num_pixels_time(self, time):
        try:
            idx = (time - 1).int()
        except AttributeError:
            idx = torch.as_tensor(time - 1).int()
        ratio, counts = self._table(idx.device)
        if self.args.select_degrade_pixel == "indexing":
            table = counts if counts is not None else ratio
        elif self.args.select_degrade_pixel == "thresholding":
            table = ratio
        else:
            raise ValueError(self.args.select_degrade_pixel)
        return torch.index_select(table, 0, idx)

    # ------------------------------------------------------------------------------------------
    # masks (one byte per pixel on the device) -- scheduler.py:278-296, 430-448
    # ------------------------------------------------------------------------------------------
    def _mask_channels(self):
        if self.args.select_degrade_pixel == "thresholding" and self.args.degrade_channel == "3-channel":
            return 3
        return 1

    def make_mask_bytes(self, black_area_num, device, out=None):
        """uint8 [B, mask_ch*HW]; consumes the CPU-generator stream exactly as the reference."""
        B = len(black_area_num)
        rng = self._rng_for(device)
        mode = self.args.select_degrade_pixel
        if mode == "indexing":
            if black_area_num.dtype.is_floating_point:
                # the reference slices with a float tensor -> TypeError (quirk q4)
                raise TypeError("select_degrade_pixel='indexing' needs integer pixel counts (use ddpm_schedule='log')")
            count = black_area_num.to(device=device, dtype=torch.int64)
            ws = self._buf("fy_words", (B * (self.image_size - 1),), torch.int32, device)
            return rng.randperm_mask(count, B, self.image_size, out=out, words_ws=ws)
        if mode == "thresholding":
            ratio = black_area_num.to(device=device, dtype=torch.float64)
            return rng.threshold_mask(ratio, B, self._mask_channels() * self.image_size, out=out)
        raise ValueError(mode)

    @staticmethod
    def _fill_mode(mean_option, mean_area):
        try:
            return _lib.FILL_CONST, float(mean_option), _lib.AREA_IMAGE
        except (ValueError, TypeError):
            pass
        area = _lib.AREA_IMAGE if mean_area == "image-wise" else _lib.AREA_CHANNEL
        if mean_option == "degraded_area":
            if mean_area not in ("image-wise", "channel-wise"):
                raise UnboundLocalError("mean_pixel")   # reference: variable never assigned
            return _lib.FILL_DEGRADED_AREA, 0.0, area
        if mean_option == "non_degraded_area":
            return _lib.FILL_NON_DEGRADED, 0.0, _lib.AREA_CHANNEL
        raise UnboundLocalError("mean_pixel")

    def _composite(self, img, mask_bytes, mask_ch, mean_option, mean_area, want_mask=True,
                   want_degrade_mask=False):
        require_cuda(img, "img")
        img = img.contiguous()
        B, C = img.shape[0], img.shape[1]
        hw = img.shape[2] * img.shape[3]
        if img.dtype == torch.float32:
            dt = _lib.MDM_F32
        elif img.dtype == torch.bfloat16:
            dt = _lib.MDM_BF16
        elif img.dtype == torch.uint8:
            dt = _lib.MDM_U8           # raw image bytes: ToTensor + Normalize(0.5, 0.5) fused into the read (utils/mydataset.py:81)
        else:
            raise TypeError(f"img dtype {img.dtype} not supported (float32 / bfloat16 / uint8)")
        if mask_ch not in (1, C):
            raise RuntimeError(f"mask with {mask_ch} channels cannot broadcast over {C} image channels")
        mode, const, area = self._fill_mode(mean_option, mean_area)
        dev = img.device
        x_t = torch.empty(B, C, img.shape[2], img.shape[3], dtype=torch.float32, device=dev)
        mask_f = torch.empty(B, mask_ch, img.shape[2], img.shape[3], dtype=torch.float32, device=dev) if want_mask else None
        dmask = torch.empty_like(x_t) if want_degrade_mask else None
        fill = torch.empty(B, C, 1, 1, dtype=torch.float32, device=dev)
        ws = self._buf("degrade_ws", (max(1, lib().mdm_degrade_ws_floats(B, C, hw)),), torch.float32, dev)
        if dt == _lib.MDM_U8:
            # the normalised fp32 image is a by-product (the trainers' loss reads it): `self.x0_normalised`
            self.x0_normalised = torch.empty_like(x_t)
            check(lib().mdm_degrade_u8(ptr(img), ptr(mask_bytes), mask_ch, mode, const, area, ptr(x_t), ptr(self.x0_normalised),
                                       ptr(mask_f), ptr(dmask), ptr(fill), ptr(ws), B, C, hw, stream_ptr(dev)))
            return x_t, mask_f, dmask, fill
        check(lib().mdm_degrade(ptr(img), dt, ptr(mask_bytes), mask_ch, mode, const, area, ptr(x_t), ptr(mask_f),
                                ptr(dmask), ptr(fill), ptr(ws), B, C, hw, stream_ptr(dev)))
        return x_t, mask_f, dmask, fill

    # ------------------------------------------------------------------------------------------
    # scheduler.py:266-323
    # ------------------------------------------------------------------------------------------
    def degrade_training(self, black_area_num, img, mean_option=None, mean_area=None, want_degrade_mask=True):
        """`want_degrade_mask=False` (not a reference argument) skips materialising `degrade_mask`, which
        only the reference's image grids read; the slot is then None."""
        mask_ch = self._mask_channels()
        mb = self.make_mask_bytes(black_area_num, img.device)
        x_t, mask_f, dmask, fill = self._composite(img, mb, mask_ch, mean_option, mean_area,
                                                   want_mask=True, want_degrade_mask=want_degrade_mask)
        masks = mask_f.expand_as(img) if mask_ch == 1 else mask_f
        mean_mask = fill.expand(img.shape[0], img.shape[1], self.height, self.width)
        return x_t, masks, dmask, mean_mask

    # scheduler.py:418-477
    def degrade_independent_base_sampling(self, black_area_num_t, img, mean_option=None, mean_area=None):
        mask_ch = self._mask_channels()
        mb = self.make_mask_bytes(black_area_num_t, img.device)
        x_t, mask_f, _, fill = self._composite(img, mb, mask_ch, mean_option, mean_area)
        masks = mask_f.expand_as(img) if mask_ch == 1 else mask_f
        mean_mask = fill.expand(img.shape[0], masks.shape[1], self.height, self.width)
        return x_t, masks, mean_mask

    # scheduler.py:480-549
    def degrade_dependent_base_sampling(self, black_area_num_t, black_area_num_next_t, img, mean_option, mean_area):
        if self.args.select_degrade_pixel != "thresholding":
            raise UnboundLocalError("masks_t")      # reference: `pass` for indexing, then uses masks_t
        mask_ch = self._mask_channels()
        B = img.shape[0]
        rng = self._rng_for(img.device)
        r_t = black_area_num_t.to(device=img.device, dtype=torch.float64)
        r_n = black_area_num_next_t.to(device=img.device, dtype=torch.float64)
        mb_t, mb_n = rng.threshold_mask(r_t, B, mask_ch * self.image_size, ratio2=r_n)
        out = []
        for mb in (mb_t, mb_n):
            x_t, mask_f, _, fill = self._composite(img, mb, mask_ch, mean_option, mean_area)
            masks = mask_f.expand_as(img) if mask_ch == 1 else mask_f
            out += [x_t, masks, fill.expand(B, masks.shape[1], self.height, self.width)]
        return tuple(out)

    # scheduler.py:572-598
    def degrade_with_mask(self, img, masks, mean_option, mean_area):
        require_cuda(img, "img")
        B, C = img.shape[0], img.shape[1]
        m = masks
        if m.dim() == 4 and m.shape[1] == C and C > 1 and m.stride(1) == 0:
            m = m[:, :1]                                      # an expanded 1-channel mask
        mask_ch = m.shape[1]
        mb = (m != 0).to(torch.uint8).contiguous().view(B, -1)
        x_t, _, _, _ = self._composite(img, mb, mask_ch, mean_option, mean_area, want_mask=False)
        return x_t

    # ------------------------------------------------------------------------------------------
    # shift noise -- scheduler.py:612-732
    # ------------------------------------------------------------------------------------------
    def get_schedule_shift_time(self, timesteps: torch.Tensor, binarymasks: torch.Tensor) -> torch.Tensor:
        require_cuda(timesteps, "timesteps")
        dev = timesteps.device
        a = self.args
        idx = timesteps.int().long() - 1
        B = len(idx)
        H, W = self.height, self.width
        ratio = torch.index_select(self._table(dev)[0], 0, idx)            # float64 [B]
        rng = self._rng_for(dev)
        kind = a.shift_type
        if kind == "1-d_constant":
            u = rng.uniform(B, -1.0, 1.0)
            shift = (u * ratio).to(a.weight_dtype)[:, None, None, None]
        elif kind == "3-d_constant":
            u = rng.uniform(B * 3, -1.0, 1.0).view(B, 3, 1, 1)
            shift = (u * ratio[:, None, None, None]).to(a.weight_dtype)
        elif kind in ("noise_reduction", "noise_with_perturbation"):
            ch = 1 if kind == "noise_reduction" else 3
            if kind == "noise_with_perturbation":
                rng.skip(B)        # the uniform perturbation is drawn, then overwritten (quirk q8)
            if B == W:
                # quirk q7: `random * ratio` broadcasts the (B,) ratio over the LAST axis when
                # B == W (per-column scaling).  Reproduced literally.
                n = rng.normal(B, ch * H * W, a.noise_mean, 1.0).view(B, ch, H, W)
                shift = (n * ratio).to(torch.float32)
            else:
                shift = rng.normal(B, ch * H * W, a.noise_mean, 1.0, ratio=ratio).view(B, ch, H, W)
        elif kind == "non_shift":
            shift = torch.zeros(B, 3, H, W, device=dev)
        elif kind == "noise_std_reduction":
            # scheduler.py:686-694: one FloatTensor(1,3,H,W).normal_(mean, std = ratio[i]) per sample.  3*H*W is a
            # multiple of 16, so the B calls consume the stream exactly like ONE 16-wide Box-Muller pass over
            # B*3*H*W uniforms; normal_ computes (radius * cos) * float(std) + mean in fp32: take the unit normals
            # from the device generator (x * 1 + 0 is exact) and apply the per-sample std and the mean as two
            # separately rounded fp32 operations.
            if (3 * H * W) % 16 != 0:
                raise NotImplementedError("noise_std_reduction: 3*H*W must be a multiple of 16 on the B200 path")
            z = rng.normal(B, 3 * H * W, 0.0, 1.0).view(B, 3, H, W)
            shift = z * ratio.to(torch.float32)[:, None, None, None]
            shift = shift + float(a.noise_mean)
        else:
            raise UnboundLocalError("shift_time")
        return shift.to(a.weight_dtype).expand_as(binarymasks)

    # scheduler.py:757-777
    def perturb_shift(self, data: torch.Tensor, shift: torch.Tensor):
        try:
            return data + shift.to(data.device)
        except RuntimeError:
            return data + shift[:, None, None, None].to(data.device)

    def perturb_shift_inverse(self, data: torch.Tensor, shift: torch.Tensor):
        try:
            return data - shift.to(data.device)
        except RuntimeError:
            return data - shift[:, None, None, None].to(data.device)

    # scheduler.py:780-794
    def get_weight_timesteps(self, timesteps: torch.Tensor, power_base=2.0):
        key = ("weight_t", float(power_base), self.updated_ddpm_num_steps, str(timesteps.device))
        power = self._ws.get(key)
        if power is None:       # cached on the device (the reference rebuilds + uploads it every call)
            alpha = torch.linspace(start=1, end=0, steps=self.updated_ddpm_num_steps)
            power = self._ws[key] = torch.pow(power_base, alpha).to(timesteps.device)
        return power[timesteps]
