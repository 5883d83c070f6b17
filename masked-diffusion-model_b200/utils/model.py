"""Drop-in for the reference's `utils/model.py` (`MyModel`, /root/reference/code/utils/model.py:3-33):
the `diffusers.UNet2DModel` configuration (block_out_channels (128,128,256,256,512,512),
layers_per_block 2, 1-5 attention stages) built on the B200 denoiser instead of diffusers."""
from mdm_b200.denoiser import UNet2DModelB200, default_config


def MyModel(dim_channel: int, dim_height: int, dim_width: int, num_attention: int = 1, device="cuda", base: int = 128):
    if dim_height != dim_width:
        raise RuntimeError("MyModel: square images only on the B200 path")
    return UNet2DModelB200(device=device, **default_config(dim_channel, dim_height, num_attention, base=base))
