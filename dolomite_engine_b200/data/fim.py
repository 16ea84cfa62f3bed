"""Fill-in-the-middle augmentation of pretraining samples (reference: data/megatron/gpt_dataset.py:162-232 and
`permute` :513-580).  Every document segment of a sample (split at the end-of-document id) is, with probability
`rate`, cut at two uniformly drawn CHARACTER positions of its decoded text into prefix / middle / suffix and re-emitted as

    PSM:  <fim_prefix> prefix <fim_suffix> suffix <fim_middle> middle
    SPM:  <fim_prefix> <fim_suffix> suffix <fim_middle> prefix middle          (with probability `spm_rate`)

then the sample is cut / padded (pad id) back to its original length.  The random stream is one
`numpy.random.RandomState(seed)` per dataset consumed in the reference's order (binomial(rate), randint x2,
binomial(spm_rate) per non-skipped segment), so the same seed reproduces the reference's samples bit for bit.

The tokenizer is duck-typed exactly as the reference uses it: `.detokenize(ids) -> str`, `.tokenize(str) -> ids`,
`.eod`, `.convert_tokens_to_ids(token)`.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

FIM_PREFIX, FIM_MIDDLE, FIM_SUFFIX, FIM_PAD = "<fim_prefix>", "<fim_middle>", "<fim_suffix>", "<fim_pad>"


@dataclass
class FIMSpec:
    rate: float
    spm_rate: float
    tokenizer: object
    prefix_id: int
    middle_id: int
    suffix_id: int
    pad_id: int
    eod_id: int

    @classmethod
    def from_tokenizer(cls, tokenizer, rate: float, spm_rate: float) -> "FIMSpec":
        if not 0 <= rate <= 1:
            raise ValueError("FIM rate must be a probability 0 <= rate <= 1")
        ids = {t: tokenizer.convert_tokens_to_ids(t) for t in (FIM_PREFIX, FIM_MIDDLE, FIM_SUFFIX, FIM_PAD)}
        return cls(rate, spm_rate, tokenizer, ids[FIM_PREFIX], ids[FIM_MIDDLE], ids[FIM_SUFFIX], ids[FIM_PAD], tokenizer.eod)


def _ids(x) -> np.ndarray:
    return np.asarray(list(x), dtype=np.int64)


def rearrange_segment(segment: np.ndarray, rng: np.random.RandomState, spec: FIMSpec) -> np.ndarray:
    """one document segment -> itself or its PSM / SPM rearrangement (length grows by 3 sentinels, re-tokenisation may
    change it further; the caller restores the sample length)"""
    if not rng.binomial(1, spec.rate):
        return segment
    text = spec.tokenizer.detokenize(segment)
    lo, hi = sorted(rng.randint(low=0, high=len(text) + 1, size=2))
    tok = spec.tokenizer.tokenize
    prefix, middle, suffix = _ids(tok(text[:lo])), _ids(tok(text[lo:hi])), _ids(tok(text[hi:]))
    if rng.binomial(1, spec.spm_rate):
        pieces = ([spec.prefix_id, spec.suffix_id], suffix, [spec.middle_id], prefix, middle)
    else:
        pieces = ([spec.prefix_id], prefix, [spec.suffix_id], suffix, [spec.middle_id], middle)
    return np.concatenate([np.asarray(p, dtype=np.int64) for p in pieces])


def apply_fim(sample: np.ndarray, rng: np.random.RandomState, spec: FIMSpec) -> np.ndarray:
    """whole sample: per-document rearrangement, end-of-document ids kept in place, original length restored"""
    n = sample.shape[0]
    breaks = np.flatnonzero(sample == spec.eod_id)
    if breaks.size == 0:
        out = rearrange_segment(sample, rng, spec)
    else:
        pieces, start = [], 0
        for b in breaks:
            b = int(b)
            if b > start:  # empty segments are skipped together with their end-of-document id, as in the reference
                pieces += [rearrange_segment(sample[start:b], rng, spec), np.asarray([spec.eod_id], dtype=np.int64)]
            start = b + 1
        pieces.append(rearrange_segment(sample[start:], rng, spec))  # the (possibly empty) tail always draws
        out = np.concatenate(pieces)
    if out.shape[0] > n:
        out = out[:n]
    elif out.shape[0] < n:
        out = np.concatenate([out, np.full(n - out.shape[0], spec.pad_id, dtype=np.int64)])
    return out.astype(np.int64, copy=False)


class HFTokenizerCodec:
    """adapts a HuggingFace tokenizer to the duck type above (decode / encode without special tokens, eos as eod)"""

    def __init__(self, tokenizer):
        self.tk = tokenizer
        self.eod = tokenizer.eos_token_id

    def detokenize(self, ids) -> str:
        return self.tk.decode([int(i) for i in ids])

    def tokenize(self, text: str) -> list[int]:
        return self.tk.encode(text, add_special_tokens=False) if text else []

    def convert_tokens_to_ids(self, token: str) -> int:
        return self.tk.convert_tokens_to_ids(token)
