"""Fill-in-the-middle augmentation of the data feed against samples the REFERENCE's GPTDataset produced
(tests/golden/fim_feed.npz, written by oracle/pin_fim.py: gpt_dataset.py:162-232, :513-580)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from dolomite_engine_b200.data import (FIMSpec, GPTDataset, MegatronBatchSampler, MMapIndexedDataset, PackedBatchLoader,
                                       apply_fim)
from pin_fim import SENTINELS, VOCAB, ToyTokenizer  # the toy tokenizer the fixture was made with (no reference import)

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "fim_feed.npz"))
PREFIX = os.path.join(ROOT, "tests", "golden", "fim_corpus")


def dataset(ci):
    lo, hi, num_samples, S, seed, n = (int(x) for x in GOLD[f"case{ci}_meta"])
    rate, spm = (float(x) for x in GOLD[f"case{ci}_rates"])
    ids = MMapIndexedDataset(PREFIX)
    spec = FIMSpec.from_tokenizer(ToyTokenizer(), rate, spm)
    return GPTDataset(ids, np.arange(lo, hi, dtype=np.int32), num_samples, S, seed, fim=spec), n, S


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_fim_samples_match_reference(ci):
    ds, n, S = dataset(ci)
    got = np.stack([ds[i]["text"] for i in range(n)])
    want = GOLD[f"case{ci}_samples"]
    assert got.dtype == np.int64 and got.shape == want.shape
    assert np.array_equal(got, want)
    plain = GPTDataset(ds.indexed_dataset, ds.indexed_indices, ds.num_samples, S, ds.random_seed)
    changed = sum(not np.array_equal(plain[i]["text"], want[i]) for i in range(n))
    assert changed > n // 4  # the fixture does exercise the transformation


def test_fim_through_the_batch_loader():
    """the native gather path applies the same rewrite, row by row in sampler order"""
    ds, n, S = dataset(0)
    mbs = 4
    loader = PackedBatchLoader(ds, MegatronBatchSampler(len(ds), 0, mbs, 1, 0), S, pin=False)
    rows = []
    for batch in loader:
        rows.append(batch["text"].numpy())
        if len(rows) * mbs >= n:
            break
    got = np.concatenate(rows)[:n]
    assert np.array_equal(got, GOLD["case0_samples"][: got.shape[0]])


def test_fim_properties():
    spec = FIMSpec.from_tokenizer(ToyTokenizer(), 1.0, 0.5)
    rng = np.random.RandomState(3)
    eod = ToyTokenizer.eod
    for trial in range(50):
        n = int(rng.randint(1, 80))
        sample = rng.randint(0, VOCAB, size=n).astype(np.int64)
        out = apply_fim(sample.copy(), rng, spec)
        assert out.shape == sample.shape and out.dtype == np.int64
        # sentinels appear in PSM / SPM order in the first segment unless truncated away
        first = out[: np.flatnonzero(out == eod)[0]] if (out == eod).any() else out
        pos = [int(np.flatnonzero(first == SENTINELS[t])[0]) for t in ("<fim_prefix>", "<fim_suffix>", "<fim_middle>")
               if (first == SENTINELS[t]).any()]
        assert pos == sorted(pos)
    # rate 0 is the identity and draws one number per segment
    ident = FIMSpec.from_tokenizer(ToyTokenizer(), 0.0, 0.5)
    s = np.arange(20, dtype=np.int64)
    assert np.array_equal(apply_fim(s, np.random.RandomState(0), ident), s)
    with pytest.raises(ValueError):
        FIMSpec.from_tokenizer(ToyTokenizer(), 1.5, 0.5)
