"""Small helpers shared by the tests."""
import numpy as np
import torch


def unpack_mask(packed, shape):
    n = int(np.prod(shape))
    return torch.from_numpy(np.unpackbits(packed)[:n].astype(np.float32)).reshape(*[int(s) for s in shape])


def torch_state_words(state_u8):
    """(uint32[624], pos) from a torch CPU generator state blob."""
    import struct
    raw = np.asarray(state_u8, dtype=np.uint8).tobytes()
    seed, left, seeded, nxt = struct.unpack_from("<QiiQ", raw, 0)
    key = np.frombuffer(raw, dtype="<u8", count=624, offset=24).astype(np.uint32)
    return key, (624 if left == 1 else int(nxt))
