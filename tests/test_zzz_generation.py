"""Decoding front end (hf_models/generation.py): the logits filters against HuggingFace's warpers on CPU; greedy / sampled
decoding against step-by-step forward passes on the GPU."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dolomite_engine_b200.hf_models.generation import _filter_logits


def test_logit_filters_match_huggingface_warpers():
    from transformers.generation.logits_process import TemperatureLogitsWarper, TopKLogitsWarper, TopPLogitsWarper

    g = torch.Generator().manual_seed(0)
    logits = torch.randn(5, 97, generator=g) * 3
    ids = torch.zeros(5, 1, dtype=torch.long)
    for temperature, top_k, top_p in [(0.7, None, None), (None, 10, None), (None, None, 0.9), (1.3, 20, 0.8), (None, 1, None),
                                      (None, 500, 0.999)]:
        want = logits.clone()
        if temperature is not None:
            want = TemperatureLogitsWarper(temperature)(ids, want)
        if top_k is not None:
            want = TopKLogitsWarper(top_k)(ids, want)
        if top_p is not None:
            want = TopPLogitsWarper(top_p)(ids, want)
        got = _filter_logits(logits.clone(), temperature, top_k, top_p)
        assert torch.equal(torch.isinf(got), torch.isinf(want)), (temperature, top_k, top_p)
        keep = ~torch.isinf(want)
        assert torch.allclose(got[keep], want[keep])


CFG = dict(vocab_size=512, n_positions=256, n_embd=256, n_layer=2, n_head=4, n_inner=512, attention_head_type="gqa",
           num_key_value_heads=2, position_embedding_type="rope", activation_function="swiglu", normalization_function="rmsnorm",
           add_bias=False, resid_pdrop=0, embd_pdrop=0, attn_pdrop=0, eos_token_id=3, bos_token_id=3, pad_token_id=3)


def _model(padding_free: bool):
    from dolomite_engine_b200.hf_models import GPTDolomiteConfig, GPTDolomiteForCausalLM

    return GPTDolomiteForCausalLM(GPTDolomiteConfig(**CFG), device=torch.device("cuda", 0),
                                  use_padding_free_transformer=padding_free)


def _prompts():
    g = torch.Generator().manual_seed(5)
    lens = [9, 4, 13, 1]
    L = max(lens)
    ids = torch.full((len(lens), L), 3, dtype=torch.long)
    mask = torch.zeros(len(lens), L, dtype=torch.long)
    for r, n in enumerate(lens):  # left padded, like the reference's inference collate (data/utils.py)
        ids[r, L - n :] = torch.randint(4, 512, (n,), generator=g)
        mask[r, L - n :] = 1
    return ids, mask


@pytest.mark.gpu
@pytest.mark.parametrize("padding_free", [False, True])
def test_greedy_generation_equals_stepwise_argmax(padding_free):
    model = _model(padding_free)
    ids, mask = _prompts()
    out = model.generate(input_ids=ids, attention_mask=mask, max_new_tokens=6, eos_token_id=-1)  # never stops early
    assert out.shape == (4, ids.shape[1] + 6) and torch.equal(out[:, : ids.shape[1]].cpu(), ids)
    # replay: every generated token is the argmax of an independent forward over the row's prefix alone
    padded = _model(False)
    padded.load_state_dict(model.state_dict())
    for r in range(4):
        n0 = int(mask[r].sum())
        row = out[r, ids.shape[1] - n0 :].cpu()
        for t in range(6):
            prefix = row[: n0 + t][None]
            with torch.no_grad():
                logits = padded(input_ids=prefix, attention_mask=torch.ones_like(prefix)).logits[0, -1].float()
            top2 = logits.topk(2).values
            if float(top2[0] - top2[1]) < 0.05:  # near tie: the batch composition may legitimately flip a bf16 argmax
                break
            assert int(logits.argmax()) == int(row[n0 + t]), (r, t)


@pytest.mark.gpu
@pytest.mark.parametrize("ng,g,hd", [(2, 2, 64), (4, 1, 80), (2, 4, 128), (1, 4, 32)])
def test_single_query_cache_attention_matches_softmax(ng, g, hd):
    """csrc/attention_decode.cu against an fp64 softmax over the cached keys: ragged cache lengths (1 .. several 128-key chunks),
    MHA / GQA / MQA slot layouts (attention/sdpa.py:11-83 with one query token)"""
    from dolomite_engine_b200 import kernels as K

    gen = torch.Generator(device="cuda").manual_seed(3)
    B, L_max = 5, 700
    lens = torch.tensor([1, 127, 128, 129, 650], dtype=torch.int32, device="cuda")
    qkv = torch.randn(B, ng * (g + 2) * hd, device="cuda", generator=gen).bfloat16()
    kc = torch.randn(B, L_max, ng * hd, device="cuda", generator=gen).bfloat16()
    vc = torch.randn(B, L_max, ng * hd, device="cuda", generator=gen).bfloat16()
    scale = hd**-0.5
    out = K.attn_decode(qkv, kc, vc, lens, ng, g, hd, scale).double().view(B, ng, g, hd)
    q = qkv.view(B, ng, g + 2, hd)[:, :, :g].double()
    for b in range(B):
        n = int(lens[b])
        k = kc[b, :n].view(n, ng, hd).double()
        v = vc[b, :n].view(n, ng, hd).double()
        p = torch.softmax(torch.einsum("ngd,lnd->ngl", q[b], k) * scale, dim=-1)
        ref = torch.einsum("ngl,lnd->ngd", p, v)
        assert torch.allclose(out[b], ref, atol=2e-2, rtol=2e-2), b


@pytest.mark.gpu
def test_cached_decoding_equals_recomputing_the_prefix():
    """generate(use_cache=True) -- one packed prefill + one decode step per token -- against use_cache=False (the packed
    forward over the whole prefix for every token): same greedy tokens up to bf16 near-ties, logits of a decode step close to
    the logits of the recomputed prefix"""
    from dolomite_engine_b200.hf_models.generation import _prefill, last_token_logits

    model = _model(True)
    ids, mask = _prompts()
    ids, mask = ids.cuda(), mask.cuda().bool()
    ref0 = last_token_logits(model, ids, mask)
    got0, cache = _prefill(model, ids, mask, 8)
    assert torch.equal(cache.lens.cpu(), mask.sum(1).int().cpu())
    assert torch.allclose(got0, ref0, atol=2e-2, rtol=2e-2)
    nxt = ref0.argmax(-1)
    ids1 = torch.cat([ids, nxt[:, None]], dim=1)
    mask1 = torch.cat([mask, torch.ones_like(mask[:, :1])], dim=1)
    ref1 = last_token_logits(model, ids1, mask1)
    got1 = model.engine.decode_step(nxt.contiguous(), cache).float()
    assert torch.equal(cache.lens.cpu(), mask1.sum(1).int().cpu())
    assert (got1 - ref1).abs().max() < 4e-2 * max(1.0, float(ref1.abs().max()))
    a = model.generate(input_ids=ids, attention_mask=mask, max_new_tokens=6, eos_token_id=-1, use_cache=True).cpu()
    b = model.generate(input_ids=ids, attention_mask=mask, max_new_tokens=6, eos_token_id=-1, use_cache=False).cpu()
    agree = (a == b).float().mean().item()
    assert agree > 0.9, agree  # a bf16 near-tie may flip one argmax and everything after it in that row


@pytest.mark.gpu
def test_generation_stops_at_eos_and_pads():
    model = _model(False)
    ids, mask = _prompts()
    free = model.generate(input_ids=ids, attention_mask=mask, max_new_tokens=5, eos_token_id=-1)
    free_gen = free[:, ids.shape[1] :].cpu()
    eos = int(free_gen[0, 0])  # make row 0's first new token the stop token
    out = model.generate(input_ids=ids, attention_mask=mask, max_new_tokens=5, eos_token_id=eos, pad_token_id=0)
    gen = out[:, ids.shape[1] :].cpu()
    # the first step is the same computation in both runs; row 0 stops there and is filled with the pad id
    assert torch.equal(gen[:, 0], free_gen[:, 0]) and int(gen[0, 0]) == eos
    assert bool((gen[0, 1:] == 0).all())
    for r in range(4):  # any row that emitted the stop token is padded from there on
        hit = (gen[r] == eos).nonzero()
        if len(hit):
            assert bool((gen[r, int(hit[0]) + 1 :] == 0).all())
    # sampling is reproducible under a seeded generator and stays inside top-k
    g1 = torch.Generator(device="cuda").manual_seed(1)
    g2 = torch.Generator(device="cuda").manual_seed(1)
    a = model.generate(input_ids=ids, attention_mask=mask, max_new_tokens=4, do_sample=True, top_k=5, temperature=0.8, generator=g1)
    b = model.generate(input_ids=ids, attention_mask=mask, max_new_tokens=4, do_sample=True, top_k=5, temperature=0.8, generator=g2)
    assert torch.equal(a, b)


@pytest.mark.gpu
def test_wrapper_generate_counts_tokens():
    from dolomite_engine_b200.model_wrapper import ModelWrapperForFinetuning

    w = ModelWrapperForFinetuning(pretrained_config=dict(model_type="gpt_dolomite", **CFG), device=torch.device("cuda", 0),
                                  use_padding_free_transformer=False)
    ids, mask = _prompts()
    toks, counts = w.generate({"input_ids": ids, "attention_mask": mask}, {"max_new_tokens": 4})
    assert len(toks) == 4 and len(counts) == 4 and all(1 <= c <= 5 for c in counts)


def test_generate_entry_loop_writes_jsonl(tmp_path):
    """generate.py:14-68 with a stand-in model: batching (incl. the short last batch), left-padded inference collate,
    one JSON line per example, generation kwargs without batch_size"""
    import json
    import types

    from dolomite_engine_b200.arguments import get_args_from_dict
    from dolomite_engine_b200.generate import build_datasets, generate

    data = tmp_path / "d"
    data.mkdir()
    with open(data / "a.jsonl", "w") as f:
        for i in range(5):
            f.write(json.dumps({"input": "x" * (i + 1), "output": "unused"}) + "\n")
    args = get_args_from_dict({
        "datasets": [{"class_name": "JSONLinesDataset", "data_name": "toy", "class_args": {"data_path": str(data)},
                      "input_format": "Q: __input__ A:", "max_input_tokens": 6}],
        "model_args": {"model_class": "AutoModelForCausalLM", "pretrained_config": {"model_type": "gpt_dolomite"}},
        "generation_parameters": {"batch_size": 2, "max_new_tokens": 3, "do_sample": False},
        "output_dir": str(tmp_path / "out")}, "inference")
    seen = []

    def fake_generate(batch, kwargs):
        assert "batch_size" not in kwargs and kwargs["max_new_tokens"] == 3
        ids, mask = batch["input_ids"], batch["attention_mask"]
        assert ids.shape == mask.shape and "labels" not in batch
        assert bool((mask[:, -1] == 1).all())  # left padding
        assert bool((ids[mask == 0] == 9).all())  # padded with eos
        seen.append(ids.shape[0])
        return [f"len{int(m.sum())}" for m in mask], [2] * ids.shape[0]

    model = types.SimpleNamespace(eos_token_id=9, generate=fake_generate)
    tokenize = lambda text: [ord(c) % 7 for c in text]  # noqa: E731
    ds = build_datasets(args, tokenize, 9)
    assert len(ds) == 1 and ds[0].data_name == "toy" and len(ds[0]) == 5
    generate(args, model, ds)
    assert seen == [2, 2, 1]
    lines = [json.loads(x) for x in open(tmp_path / "out" / "output-toy.jsonl")]
    assert len(lines) == 5 and lines[0] == {"generated_text": "len6", "num_generated_tokens": 2}  # truncated to 6 prompt tokens
    assert os.path.isfile(tmp_path / "out" / "inference_config.yml")
    with pytest.raises(Exception, match="model_args need to be specified"):
        get_args_from_dict({"datasets": args.model_dump()["datasets"], "generation_parameters": {"batch_size": 1, "max_new_tokens": 1},
                            "output_dir": "x"}, "inference")  # neither model_args nor load_args
