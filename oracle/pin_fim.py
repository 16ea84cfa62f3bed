"""Pins data/fim.py against the reference's GPTDataset with fim_rate != 0 (runs only where /root/reference exists) and
writes tests/golden/fim_corpus.{bin,idx} (small vocabulary, every document ends with the end-of-document id 17, which
also occurs inside documents and doubled) and tests/golden/fim_feed.npz: samples the REFERENCE produced from it with a toy
character-level tokenizer (token id t <-> chr(0x100 + t); sentinels / pad are ids above the corpus vocabulary).

    python oracle/pin_fim.py

Test infrastructure only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from pin_data_feed import OUT as FEED, compile_reference_helpers, import_reference_data_modules  # noqa: E402

VOCAB = 60
SENTINELS = {"<fim_prefix>": VOCAB + 1, "<fim_middle>": VOCAB + 2, "<fim_suffix>": VOCAB + 3, "<fim_pad>": VOCAB + 4}


class ToyTokenizer:
    """one character per token id, except that the two-character string 'ab' style merges are emulated by mapping ids
    divisible by 7 to TWO characters -- so character boundaries can fall inside a token and re-tokenisation changes
    lengths, like a BPE vocabulary"""

    eod = 17

    def detokenize(self, ids):
        return "".join(chr(0x100 + int(t)) + ("~" if int(t) % 7 == 0 else "") for t in ids)

    def tokenize(self, text):
        out, i = [], 0
        while i < len(text):
            c = text[i]
            if c == "~":  # an orphaned second half (boundary fell inside a token)
                out.append(VOCAB)
                i += 1
                continue
            t = ord(c) - 0x100
            if t % 7 == 0 and i + 1 < len(text) and text[i + 1] == "~":
                i += 1
            out.append(t)
            i += 1
        return out

    def convert_tokens_to_ids(self, tok):
        return SENTINELS[tok]


def main():
    compile_reference_helpers()
    m = import_reference_data_modules()
    Cfg = m["blended_megatron_dataset_config"].GPTDatasetConfig
    Split = sys.modules["dolomite_engine.data.megatron.utils"].Split
    import torch

    prefix = os.path.join(ROOT, "tests", "golden", "fim_corpus")
    rng = np.random.default_rng(11)
    b = m["indexed_dataset"].MMapIndexedDatasetBuilder(prefix + ".bin", dtype=np.uint16)
    for d in range(37):
        toks = np.append(rng.integers(0, VOCAB, size=int(rng.integers(1, 70))), ToyTokenizer.eod)
        if d % 9 == 4:
            toks = np.append(toks, ToyTokenizer.eod)  # doubled end-of-document: an empty segment
        b.add_item(torch.from_numpy(toks.astype(np.int64)))
        b.end_document()
    b.finalize(prefix + ".idx")
    ids = m["indexed_dataset"].MMapIndexedDataset(prefix)
    out = {}
    cases = [(0, 37, 60, 24, 1234, 0.7, 0.5), (0, 30, 45, 64, 9, 1.0, 0.0), (5, 37, 50, 16, 77, 0.4, 1.0)]
    for ci, (lo, hi, num_samples, S, seed, rate, spm) in enumerate(cases):
        cfg = Cfg(is_built_on_rank=True, random_seed=seed, sequence_length=S, blend=[prefix],
                  split="100,0,0", path_to_cache=f"/tmp/pin_fim_cache_{ci}", return_document_ids=False, fim_rate=rate,
                  fim_spm_rate=spm)
        ds = m["gpt_dataset"].GPTDataset(ids, np.arange(lo, hi, dtype=np.int32), num_samples, Split.train, ToyTokenizer(),
                                         cfg, True)
        n = min(len(ds), 48)
        out[f"case{ci}_meta"] = np.asarray([lo, hi, num_samples, S, seed, n], dtype=np.int64)
        out[f"case{ci}_rates"] = np.asarray([rate, spm])
        out[f"case{ci}_samples"] = np.stack([ds[i]["text"] for i in range(n)])  # sequential: one shared random stream
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fim_feed.npz"), **out)
    print("wrote tests/golden/fim_feed.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
