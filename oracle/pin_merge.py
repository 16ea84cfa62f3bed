"""Pins `MMapIndexedDatasetBuilder.add_index` / data/merge.py against the reference's builder (runs only where
/root/reference exists): merges tests/golden/data_feed/corpus_a with itself and with tests/golden/fim_corpus (both uint16)
using the REFERENCE's `add_index` and records the sha256 of the resulting .bin / .idx in tests/golden/merge_expected.json.

    python oracle/pin_merge.py

Test infrastructure only."""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from pin_data_feed import compile_reference_helpers, import_reference_data_modules  # noqa: E402

CASES = {"a_a": ["data_feed/corpus_a", "data_feed/corpus_a"], "a_fim_a": ["data_feed/corpus_a", "fim_corpus", "data_feed/corpus_a"]}


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    compile_reference_helpers()
    m = import_reference_data_modules()
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for name, parts in CASES.items():
            b = m["indexed_dataset"].MMapIndexedDatasetBuilder(os.path.join(d, name + ".bin"), dtype=np.uint16)
            for p in parts:
                b.add_index(os.path.join(ROOT, "tests", "golden", p))
            b.finalize(os.path.join(d, name + ".idx"))
            ds = m["indexed_dataset"].MMapIndexedDataset(os.path.join(d, name))
            out[name] = {"parts": parts, "bin": sha(os.path.join(d, name + ".bin")), "idx": sha(os.path.join(d, name + ".idx")),
                         "sequences": len(ds), "documents": int(ds.document_indices.shape[0])}
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "merge_expected.json"), "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main()
