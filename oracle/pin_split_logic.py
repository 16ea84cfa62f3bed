"""Pins the train/val/test split arithmetic of the data feed against the reference (runs only where /root/reference exists):
`_parse_and_normalize_split` (blended_megatron_dataset_config.py:98-120) and `_get_split_indices`
(blended_megatron_dataset_builder.py:376-397) are evaluated on a grid and stored in tests/golden/split_logic.json.

    python oracle/pin_split_logic.py

Test infrastructure only."""
import json
import os
import sys
import typing

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from pin_data_feed import compile_reference_helpers, import_reference_data_modules  # noqa: E402

SPLITS = ["100,0,0", "969,30,1", "98,2,0", "0.8,0.1,0.1", "1,1,1", "90,10", "100", "949,50,1", "3,3,4", "0,100,0", "98/2/0"]
SIZES = [1, 7, 10, 37, 1000, 12345]


def main():
    compile_reference_helpers()
    cfg = import_reference_data_modules()["blended_megatron_dataset_config"]
    src = open("/root/reference/dolomite_engine/data/megatron/blended_megatron_dataset_builder.py").read()
    # the builder module drags torch.distributed machinery in; only the pure function is evaluated
    ns = {"List": typing.List}
    exec(src[src.index("def _get_split_indices"):src.index("def _get_prefixes_weights_and_sizes_for_blend")], ns)
    out = []
    for s in SPLITS:
        vec = cfg._parse_and_normalize_split(s)
        out.append({"split": s, "vector": list(vec), "bounds": {str(n): list(ns["_get_split_indices"](vec, n)) for n in SIZES}})
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "split_logic.json"), "w"), indent=1)
    print("pinned", len(out), "splits x", len(SIZES), "sizes")


if __name__ == "__main__":
    main()
