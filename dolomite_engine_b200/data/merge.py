"""Merge Megatron `.bin/.idx` stores (reference tool: tools/megatron_dataset/merge_data.py): every store found under
`--input-prefixes` (or every `<name>.idx` with a matching `.bin` in `--input-directory`) is appended, in sorted order, to
`<output_prefix>.{bin,idx}`.

    python -m dolomite_engine_b200.data.merge --input-directory shards/ --output-prefix merged/corpus
"""

from __future__ import annotations

import argparse
import os

from .indexed_dataset import MMapIndexedDataset, MMapIndexedDatasetBuilder, get_bin_path, get_idx_path


def merge(prefixes: list[str], output_prefix: str) -> int:
    """-> number of sequences written"""
    if not prefixes:
        raise ValueError("nothing to merge")
    for p in prefixes:
        if not MMapIndexedDataset.exists(p):
            raise FileNotFoundError(f"{p}: need both {get_idx_path(p)} and {get_bin_path(p)}")
    dtype = MMapIndexedDataset(prefixes[0]).dtype
    os.makedirs(os.path.dirname(os.path.abspath(output_prefix)), exist_ok=True)
    builder = MMapIndexedDatasetBuilder(get_bin_path(output_prefix), dtype=dtype)
    for p in prefixes:
        builder.add_index(p)
    builder.finalize(get_idx_path(output_prefix))
    return len(builder.sequence_lengths)


def prefixes_in(directory: str) -> list[str]:
    names = sorted(f[: -len(".idx")] for f in os.listdir(directory) if f.endswith(".idx"))
    return [os.path.join(directory, n) for n in names if os.path.exists(os.path.join(directory, n + ".bin"))]


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--input-directory")
    ap.add_argument("--input-prefixes", nargs="*")
    ap.add_argument("--output-prefix", required=True)
    a = ap.parse_args()
    prefixes = list(a.input_prefixes or []) + (prefixes_in(a.input_directory) if a.input_directory else [])
    print(f"merged {merge(prefixes, a.output_prefix)} sequences of {len(prefixes)} stores into {a.output_prefix}")


if __name__ == "__main__":
    main()
