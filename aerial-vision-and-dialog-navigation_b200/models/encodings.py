"""Positional encoding of the reference (src/models/encodings.py:7-49).

``PosEncoding`` keeps the reference's ``pe`` buffer (``[1, max_len, d_model]``,
so ``state_dict`` key ``encoder_vl.enc_pos.pe`` matches); the addition itself is
fused into ``avdn_embed_fwd`` (see ``_et_engine.ETEngine.encode``).
"""
import math

import torch
from torch import nn


def sinusoid_table(d_model, max_len):
    """encodings.py:12-20, float32."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


class PosEncoding(nn.Module):
    def __init__(self, d_model, max_len=1250):
        super().__init__()
        self.d_model = d_model
        self.register_buffer("pe", sinusoid_table(d_model, max_len)[None])
