"""Pins the YAML argument schema (arguments.py) against the reference's pydantic classes (runs only where /root/reference
exists): the reference's `arguments.py` is imported read-only under stubbed parents (`peft` and `dolomite_engine.utils` are
stubs) and every section's field names and defaults are written to tests/golden/arguments_schema.json.

    python oracle/pin_arguments_schema.py

Test infrastructure only."""
import enum
import importlib.util
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/dolomite_engine"


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def plain(v):
    if isinstance(v, enum.Enum):
        return v.value
    if hasattr(type(v), "model_fields"):
        return {k: plain(getattr(v, k)) for k in type(v).model_fields}
    if isinstance(v, type):
        return v.__name__
    if isinstance(v, (list, tuple)):
        return [plain(x) for x in v]
    return v


def main():
    pkg = types.ModuleType("dolomite_engine")
    pkg.__path__ = []
    sys.modules["dolomite_engine"] = pkg
    peft = types.ModuleType("peft")
    peft.PromptTuningInit = enum.Enum("PromptTuningInit", {"TEXT": "TEXT", "RANDOM": "RANDOM"})
    sys.modules["peft"] = peft
    load("dolomite_engine.enums", f"{REF}/enums.py")
    load("dolomite_engine.defaults", f"{REF}/defaults.py")
    pyd = load("dolomite_engine.utils.pydantic", f"{REF}/utils/pydantic.py")
    utils = types.ModuleType("dolomite_engine.utils")
    utils.__path__ = []
    utils.BaseArgs = pyd.BaseArgs
    utils.load_yaml = utils.set_logger = utils.log_rank_0 = lambda *a, **k: None
    utils.run_rank_n = lambda f, *a, **k: f
    utils.normalize_dtype_string = lambda s: {"float32": "fp32", "float16": "fp16", "bfloat16": "bf16"}.get(s, s)
    sys.modules["dolomite_engine.utils"] = utils
    ref = load("dolomite_engine.arguments", f"{REF}/arguments.py")
    out = {}
    for name in dir(ref):
        cls = getattr(ref, name)
        if isinstance(cls, type) and issubclass(cls, pyd.BaseArgs) and cls is not pyd.BaseArgs:
            out[name] = {k: plain(f.default) for k, f in cls.model_fields.items()}
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "arguments_schema.json"), "w"), indent=1, sort_keys=True)
    print("pinned", len(out), "sections,", sum(len(v) for v in out.values()), "keys")


if __name__ == "__main__":
    main()
