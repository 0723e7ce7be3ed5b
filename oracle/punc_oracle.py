"""ORACLE for the CT-Transformer punctuation network — CPU fp32 restatement, TEST INFRASTRUCTURE ONLY (same rules as
paraformer_oracle.py: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this).

Restates CTTransformer.punc_forward (funasr/models/ct_transformer/model.py:112-125): Embedding -> SANMEncoder (the same
encoder restatement the Paraformer oracle uses, input_layer "pe") -> Linear -> arg-max per token.
Parity status: PINNED — tests/golden/punc_*.npz hold the unmodified reference's AutoModel(model="CTTransformer").generate()
text and punc_array for the same seeded weights and texts (oracle/make_vad_golden.py, section "punc"); tests/test_punc_host.py
checks this file and the host text logic of funasr_b200/punc.py against them.
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch
import torch.nn.functional as F

import paraformer_oracle as O

Tensor = torch.Tensor


def punc_logits(token_ids: Sequence[int], p: Dict[str, Tensor], layers: int, heads: int) -> Tensor:
    """[T] token ids -> [T, n_punc] (model.py:112-125, batch of one mini-sentence like the reference's inference loop)."""
    ids = torch.as_tensor(list(token_ids), dtype=torch.long)[None]
    x = F.embedding(ids, p["embed.weight"])
    h, _ = O.encoder(x, torch.tensor([ids.shape[1]]), p, layers, heads=heads)
    return F.linear(h, p["decoder.weight"], p["decoder.bias"])[0]


def punc_ids(token_ids: Sequence[int], p: Dict[str, Tensor], layers: int, heads: int) -> Tensor:
    """model.py:346-349: indices of the arg-max over the punctuation classes."""
    return punc_logits(token_ids, p, layers, heads).argmax(-1)
