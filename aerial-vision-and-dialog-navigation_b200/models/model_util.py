"""Hot-path subset of the reference's ``models/model_util.py``."""
import torch

from .. import _lib


def _masks(lens, len_lang, len_frames, device):
    lens_t = torch.as_tensor(list(lens), dtype=torch.int32, device=device)
    B = int(lens_t.numel())
    S = len_lang + 2 * len_frames
    mask_pad = torch.empty((B, S), dtype=torch.uint8, device=device)
    mask_attn = torch.empty((S, S), dtype=torch.float32, device=device)
    _lib.call("avdn_build_masks", _lib.ptr(lens_t), B, len_lang, len_frames, _lib.ptr(mask_pad), _lib.ptr(mask_attn))
    return mask_pad.bool(), mask_attn


def generate_attention_mask(len_lang, len_frames, device, num_input_actions=0):
    """Additive float mask ``[(L+2T),(L+2T)]`` (0 / -inf), bit-identical to the
    reference's (src/models/model_util.py:213-241).  The transformer kernels
    never read this tensor -- they evaluate the same predicate in registers --
    it exists for API parity and the bit-exact tests."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("generate_attention_mask runs on the CUDA device only (no CPU fallback)")
    return _masks([len_frames], len_lang, len_frames, dev)[1]


def generate_pad_mask(lengths, len_lang, device):
    """Key-padding mask of ``EncoderVL.forward`` (src/models/enc_vl.py:44-55): bool ``[B, L+2*Tmax]``."""
    return _masks(lengths, len_lang, int(max(lengths)), torch.device(device))[0]
