"""Pins the config surface (hf_models/config.py) against the reference's own classes (runs only where /root/reference
exists): field defaults of CommonConfig / MoEDolomiteConfig and the outcome of the constructor's consistency checks on a grid
are taken from the REFERENCE classes and stored in tests/golden/config_surface.json.  Two fields are skipped because
transformers 5.x (the base class in this image) rewrites them under the reference (`rope_scaling`, `tie_word_embeddings`).

Deliberate deviation, recorded in the fixture as `"alias": true`: when HuggingFace-style ALIAS names are passed to the
constructor (`hidden_size=...`), the reference applies them only after its derived defaults and head-count checks ran on the
canonical defaults (config.py:50-111), e.g. `CommonConfig(hidden_size=96)` keeps n_inner = 4 * 768.  This implementation
resolves aliases first; for such cases only the alias mapping itself is compared.

    python oracle/pin_config_surface.py

Test infrastructure only."""
import importlib.util
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/dolomite_engine/hf_models"
SKIP = {"rope_scaling", "tie_word_embeddings"}
GRID = [dict(attention_head_type="gqa"), dict(attention_head_type="gqa", num_key_value_heads=5),
        dict(attention_head_type="mha", num_key_value_heads=3), dict(attention_head_type="mha", n_head=6),
        dict(attention_head_type="mqa", num_key_value_heads=2), dict(attention_head_type="mqa"),
        dict(attention_head_type="gqa", num_key_value_heads=4, n_head=12), dict(attention_multiplier=0.5, scale_attn_weights=False),
        dict(attention_multiplier=0.5), dict(n_inner=None, n_embd=100), dict(n_inner=77), dict(init_method="xavier"),
        dict(init_method="mup", m_width=4.0), dict(position_embedding_type="alibi"), dict(position_embedding_type="sinusoidal"),
        dict(attention_head_type="xyz"), dict(hidden_size=96, num_attention_heads=6, num_hidden_layers=3, max_position_embeddings=77)]
PROBE = ("num_key_value_heads", "n_inner", "n_embd", "n_head", "n_layer", "n_positions", "multi_query")


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def outcome(cls, kw):
    try:
        c = cls(**kw)
    except Exception as e:  # noqa
        return {"error": type(e).__name__}
    return {k: getattr(c, k, None) for k in PROBE}


def main():
    from dolomite_engine_b200.hf_models.config import CommonConfig, MoEDolomiteConfig

    for name in ("dolomite_engine", "dolomite_engine.hf_models", "dolomite_engine.hf_models.models",
                 "dolomite_engine.hf_models.models.moe_dolomite"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    load("dolomite_engine.hf_models.enums", f"{REF}/enums.py")
    common = load("dolomite_engine.hf_models.config", f"{REF}/config.py")
    moe = load("dolomite_engine.hf_models.models.moe_dolomite.config", f"{REF}/models/moe_dolomite/config.py")
    ref_c, ref_m = common.CommonConfig(), moe.MoEDolomiteConfig()
    out = {"common_defaults": {k: getattr(ref_c, k) for k, _ in CommonConfig.all_fields() if k not in SKIP},
           "moe_defaults": {k: getattr(ref_m, k) for k, _ in MoEDolomiteConfig.all_fields() if k not in SKIP},
           "grid": [{"kwargs": kw, "alias": any(k in ("hidden_size", "num_attention_heads", "num_hidden_layers",
                                                       "max_position_embeddings") for k in kw),
                     "common": outcome(common.CommonConfig, kw), "moe": outcome(moe.MoEDolomiteConfig, kw)} for kw in GRID]}
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "config_surface.json"), "w"), indent=1)
    print("pinned", len(out["common_defaults"]), "+", len(out["moe_defaults"]), "defaults and", len(GRID), "constructor cases")


if __name__ == "__main__":
    main()
