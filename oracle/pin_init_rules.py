"""Pins the parameter-initialisation rules (engine._block_specs: normal(std) per tensor, ones / zeros for norms and biases)
against the reference's own modules (runs only where /root/reference exists): `GPTDolomiteBlock` is instantiated for dense
configs (normal and muP init) and the `std` every ParameterizedLinear was constructed with is stored in
tests/golden/init_rules.json (attention/base.py:73-86, mlp.py:26-41, layer.py:33-47).

    python oracle/pin_init_rules.py

Test infrastructure only."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

CASES = {
    "normal_gqa_bias": dict(n_embd=64, n_head=4, num_key_value_heads=2, n_layer=6, n_inner=96, attention_head_type="gqa",
                            add_bias=True, activation_function="swiglu", initializer_range=0.02, init_method="normal"),
    "mup_mha": dict(n_embd=64, n_head=4, n_layer=8, n_inner=128, attention_head_type="mha", add_bias=False,
                    activation_function="swiglu", initializer_range=0.1, init_method="mup", m_width=4.0),
    "gelu_mqa": dict(n_embd=64, n_head=4, n_layer=3, n_inner=256, attention_head_type="mqa", add_bias=True,
                     activation_function="gelu_pytorch_tanh", normalization_function="layernorm", initializer_range=0.02),
}


def main():
    import dolomite_oracle as O
    from validate_against_reference import import_reference, ref_config

    R = import_reference()
    out = {}
    for name, kw in CASES.items():
        cfg = O.OracleConfig(vocab_size=264, n_positions=64, position_embedding_type="rope",
                             **{"normalization_function": "rmsnorm", **kw})
        block = R.GPTDolomiteBlock(ref_config(cfg), "torch", "eager", False, 1)
        stds = {}
        for mod_name, mod in block.named_modules():
            if hasattr(mod, "std") and hasattr(mod, "weight"):
                stds[f"{mod_name}.weight"] = float(mod.std)
        shapes = {k: list(v.shape) for k, v in block.state_dict().items()}
        out[name] = {"config": kw, "std": stds, "shapes": shapes}
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "init_rules.json"), "w"), indent=1)
    print({k: v["std"] for k, v in out.items()})


if __name__ == "__main__":
    main()
