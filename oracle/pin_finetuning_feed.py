"""Golden fixtures for the SFT collate / example construction, produced by the reference's own `collate_fn`
(data/utils.py) and `BaseDataset.get_input_output_token_ids` (data/base.py) in this container.

    python oracle/pin_finetuning_feed.py      -> tests/golden/finetuning_feed.json
"""
import enum
import importlib.util
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/dolomite_engine"


def load_reference():
    pkg = types.ModuleType("dolomite_engine")
    pkg.__path__ = []
    sys.modules["dolomite_engine"] = pkg
    enums = types.ModuleType("dolomite_engine.enums")

    class Mode(enum.Enum):
        training = "training"
        inference = "inference"

    class LossMask(enum.Enum):
        output_only = "output_only"
        no_mask = "no_mask"

    class DatasetSplit(enum.Enum):
        train = "train"
        val = "val"
        test = "test"

    enums.Mode, enums.LossMask, enums.DatasetSplit = Mode, LossMask, DatasetSplit
    sys.modules["dolomite_engine.enums"] = enums
    defaults = types.ModuleType("dolomite_engine.defaults")  # defaults.py: the two format placeholders
    defaults.INPUT_FORMAT, defaults.OUTPUT_FORMAT = "__input__", "__output__"
    sys.modules["dolomite_engine.defaults"] = defaults
    data = types.ModuleType("dolomite_engine.data")
    data.__path__ = [os.path.join(REF, "data")]
    sys.modules["dolomite_engine.data"] = data

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    utils = load("dolomite_engine.data.utils", "data/utils.py")
    base = load("dolomite_engine.data.base", "data/base.py")
    return utils, base, Mode, LossMask


class ToyTokenizer:
    """deterministic whitespace tokenizer standing in for AutoTokenizer (the reference only calls it and reads eos_token_id)"""

    eos_token_id = 2

    def __call__(self, text, add_special_tokens=False):
        return {"input_ids": [3 + (sum(map(ord, w)) % 97) for w in text.split()]}


def main():
    utils, base, Mode, LossMask = load_reference()
    tok = ToyTokenizer()
    raw = [("translate the cat sat", "le chat"), ("a", "b c d e f g h"), ("one two three four five six seven", "x"),
           ("p q", "r s")]
    out = {"raw": raw, "examples": [], "collate": []}
    for mi, mo in [(None, None), (3, 4)]:
        ds = base.BaseDataset.__new__(base.BaseDataset)
        ds.tokenizer, ds.is_encoder_decoder, ds.mode = tok, False, Mode.training
        ds.max_input_tokens, ds.max_output_tokens = mi, mo
        exs = [ds.get_input_output_token_ids(i, o) for i, o in raw]
        out["examples"].append({"max_input_tokens": mi, "max_output_tokens": mo, "examples": exs})
        for pf in (True, False):
            for lm in (LossMask.output_only, LossMask.no_mask):
                if not pf and lm == LossMask.no_mask:
                    continue  # the reference builds a ragged tensor there unless all rows have equal length
                r = utils.collate_fn(exs, Mode.training, lm, tok.eos_token_id, False, pf)
                r = {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in r.items()}
                out["collate"].append({"max_input_tokens": mi, "max_output_tokens": mo, "padding_free": pf,
                                       "loss_mask": lm.value, "result": r})
    path = os.path.join(ROOT, "tests", "golden", "finetuning_feed.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, len(out["collate"]), "collate cases")


if __name__ == "__main__":
    main()
