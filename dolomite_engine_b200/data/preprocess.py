"""Corpus preparation for the pretraining feed (reference tool: tools/megatron_dataset/preprocess_data.py): JSON-lines
text -> one Megatron `.bin/.idx` token store per JSON key, `<output_prefix>_<key>.{bin,idx}`, one document per line,
token width chosen from the vocabulary size, optional end-of-document id appended.  Empty documents are skipped, like the
reference.

    python -m dolomite_engine_b200.data.preprocess --input corpus.jsonl --tokenizer <dir> --output-prefix out/corpus \\
        --json-keys text --append-eod --workers 8 --chunk-size 64
"""

from __future__ import annotations

import argparse
import json
import multiprocessing
from typing import Callable, Iterable, Iterator

from .indexed_dataset import MMapIndexedDatasetBuilder, optimal_dtype

_ENCODE: Callable[[str], list[int]] | None = None
_KEYS: list[str] = []
_EOD: int | None = None


def _init_worker(tokenizer_path: str, keys: list[str], append_eod: bool) -> None:
    from transformers import AutoTokenizer

    tk = AutoTokenizer.from_pretrained(tokenizer_path)
    configure(tk.encode, keys, tk.eos_token_id if append_eod else None)


def configure(encode: Callable[[str], list[int]], keys: list[str], eod: int | None) -> None:
    global _ENCODE, _KEYS, _EOD
    _ENCODE, _KEYS, _EOD = encode, list(keys), eod


def encode_line(line: str) -> dict[str, list[int]]:
    """one JSON line -> {key: token ids}; keys whose text tokenises to nothing are dropped"""
    record = json.loads(line)
    out = {}
    for key in _KEYS:
        ids = list(_ENCODE(record[key]))
        if ids:
            if _EOD is not None:
                ids.append(_EOD)
            out[key] = ids
    return out


def write_stores(encoded: Iterable[dict[str, list[int]]], output_prefix: str, keys: list[str], vocab_size: int | None) -> dict[str, int]:
    """-> documents written per key"""
    dtype = optimal_dtype(vocab_size)
    builders = {k: MMapIndexedDatasetBuilder(f"{output_prefix}_{k}.bin", dtype=dtype) for k in keys}
    counts = {k: 0 for k in keys}
    for item in encoded:
        for key, ids in item.items():
            builders[key].add_item(ids)
            builders[key].end_document()
            counts[key] += 1
    for k in keys:
        builders[k].finalize(f"{output_prefix}_{k}.idx")
    return counts


def _lines(path: str) -> Iterator[str]:
    with open(path, encoding="utf-8") as f:
        for line in f:
            if line.strip():
                yield line


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--input", required=True, help="JSON-lines file")
    ap.add_argument("--json-keys", nargs="+", default=["text"])
    ap.add_argument("--tokenizer", required=True)
    ap.add_argument("--append-eod", action="store_true")
    ap.add_argument("--output-prefix", required=True)
    ap.add_argument("--workers", type=int, default=1)
    ap.add_argument("--chunk-size", type=int, default=64)
    a = ap.parse_args()
    if not a.input.endswith(".jsonl"):
        raise NotImplementedError("only JSON-lines input: HuggingFace datasets / zstd archives need packages and network this "
                                  "image does not have")
    from transformers import AutoTokenizer

    vocab = AutoTokenizer.from_pretrained(a.tokenizer).vocab_size
    if a.workers > 1:
        with multiprocessing.Pool(a.workers, initializer=_init_worker, initargs=(a.tokenizer, a.json_keys, a.append_eod)) as pool:
            counts = write_stores(pool.imap(encode_line, _lines(a.input), a.chunk_size), a.output_prefix, a.json_keys, vocab)
    else:
        _init_worker(a.tokenizer, a.json_keys, a.append_eod)
        counts = write_stores(map(encode_line, _lines(a.input)), a.output_prefix, a.json_keys, vocab)
    print(json.dumps({"documents": counts, "output_prefix": a.output_prefix}))


if __name__ == "__main__":
    main()
