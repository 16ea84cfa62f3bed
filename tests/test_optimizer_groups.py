"""CPU: muP parameter groups of the fused AdamW (reference: optimization/optimizer.py:86-126 `_get_param_groups`, method mup: the
non-bias parameters of every Attention / MLP module train with lr / m_width).  The reference builds two torch param groups; this
engine keeps one flat fp32 shard per unit and runs its AdamW kernel once per RANGE of the shard with the range's learning rate.
Checked here without a GPU: (1) the ranges tile every rank's shard and carry the multiplier of the parameter they belong to, for
any data-parallel degree; (2) the optimizer, with the kernel entry replaced by the same arithmetic in torch, follows
torch.optim.AdamW built with the reference's two groups, through a learning-rate schedule.  The kernel itself is checked on the GPU
against torch.optim.AdamW (tests/test_gpu_kernels.py::test_optimizer_kernels_vs_torch: any length, any 4-byte aligned start)."""

import math

import pytest
import torch

from dolomite_engine_b200.engine import FlatUnit, _block_specs, _root_specs
from dolomite_engine_b200.hf_models import GPTDolomiteConfig, MoEDolomiteConfig
from dolomite_engine_b200.optimization import get_optimizer, get_scheduler, lr_scale_segments, mup_lr_scale

KW = dict(n_embd=64, n_head=4, n_layer=2, n_inner=96, vocab_size=264, n_positions=32, attention_head_type="mha", add_bias=True,
          activation_function="swiglu", position_embedding_type="rope", normalization_function="rmsnorm", resid_pdrop=0,
          embd_pdrop=0, attn_pdrop=0, init_method="mup", m_width=4.0, m_emb=2.0, m_residual=0.5)


def _is_mup(name: str) -> bool:  # restated from optimizer.py:100-107: parameters of Attention / MLP modules, biases excluded
    parts = name.split(".")
    return len(parts) > 3 and parts[0] == "transformer" and parts[1] == "h" and parts[3] in ("attn", "mlp") and not name.endswith("bias")


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_learning_rate_ranges_tile_every_shard(world):
    cfg = GPTDolomiteConfig(**KW)
    scale_of = mup_lr_scale(cfg)
    for specs, name in ((_root_specs(cfg), "root"), (_block_specs(cfg, 1), "h.1")):
        units = [FlatUnit(name, specs, world, r) for r in range(world)]
        expect = torch.ones(units[0].padded)
        for s in units[0].specs:
            assert scale_of(s.name) == (0.25 if _is_mup(s.name) else 1.0), s.name
            expect[s.offset : s.offset + s.numel] = scale_of(s.name)
        got = torch.full((units[0].padded,), float("nan"))
        for u in units:
            segs = lr_scale_segments(u, scale_of)
            assert segs[0][0] == 0 and segs[-1][1] == u.shard_numel
            assert all(a[1] == b[0] for a, b in zip(segs, segs[1:]))  # contiguous ...
            assert all(a[2] != b[2] for a, b in zip(segs, segs[1:]))  # ... and merged
            for lo, hi, sc in segs:
                got[u.rank * u.shard_numel + lo : u.rank * u.shard_numel + hi] = sc
        # every element of a parameter carries its parameter's multiplier (gaps between parameters may carry either neighbour's)
        for s in units[0].specs:
            assert torch.equal(got[s.offset : s.offset + s.numel], expect[s.offset : s.offset + s.numel]), s.name
        assert not torch.isnan(got).any()


def _torch_adamw_step(p, g, m, v, pb, lr, b1, b2, eps, wd, step, clip=None):
    """the arithmetic of dolomite_b200_adamw_step (= torch.optim.AdamW) on views of the flat buffers"""
    gi = g * (1.0 if clip is None else float(clip))
    p.mul_(1 - lr * wd)
    m.mul_(b1).add_(gi, alpha=1 - b1)
    v.mul_(b2).addcmul_(gi, gi, value=1 - b2)
    denom = v.sqrt() / math.sqrt(1 - b2**step) + eps
    p.addcdiv_(m, denom, value=-lr / (1 - b1**step))
    if pb is not None:
        pb.copy_(p)


def test_fused_adamw_with_mup_groups_follows_torch_adamw_with_the_reference_groups(monkeypatch):
    from dolomite_engine_b200 import kernels as K
    from dolomite_engine_b200.distributed import ShardedDataParallel
    from dolomite_engine_b200.model_wrapper import ModelWrapperForPretraining

    calls = []

    def fake(p, g, m, v, pb, lr, b1, b2, eps, wd, step, clip=None):
        calls.append((p.numel(), lr / opt.param_groups[0]["lr"] if opt.param_groups[0]["lr"] else 1.0))
        _torch_adamw_step(p, g, m, v, pb, lr, b1, b2, eps, wd, step, clip)

    monkeypatch.setattr(K, "adamw_step", fake)
    cfg = dict(KW, model_type="gpt_dolomite")
    w = ModelWrapperForPretraining(pretrained_config=cfg, micro_batch_size=1, sequence_length=16, device=torch.device("cpu"))
    sdp = ShardedDataParallel(w, None)
    oargs = {"lr": 3e-3, "weight_decay": 0.1, "betas": [0.9, 0.95], "eps": 1e-10}
    opt = get_optimizer("DolomiteFusedAdamW", oargs, sdp, params_group_method="mup")
    sched = get_scheduler(opt, 2, 0, None, 6, "cosine", 0.1)
    # the reference's two groups over per-name copies of the same parameters
    sd = {k: v.clone() for k, v in sdp.state_dict().items()}
    ref = {k: torch.nn.Parameter(v.clone()) for k, v in sd.items()}
    groups = [{"params": [p for n, p in ref.items() if not _is_mup(n)]},
              {"params": [p for n, p in ref.items() if _is_mup(n)], "lr": oargs["lr"] / cfg["m_width"]}]
    assert len(groups[1]["params"]) == 4 * cfg["n_layer"]
    ropt = torch.optim.AdamW(groups, lr=oargs["lr"], weight_decay=0.1, betas=(0.9, 0.95), eps=1e-10)
    rsched = get_scheduler(ropt, 2, 0, None, 6, "cosine", 0.1)
    g = torch.Generator().manual_seed(0)
    eng = sdp.engine
    for step in range(5):
        for n, u, s in eng.named_views():
            grad = torch.randn(s.shape, generator=g) * 0.05
            u.gviews[n].copy_(grad)
            ref[n].grad = grad.clone()
        opt.step()
        sched.step()
        ropt.step()
        rsched.step()
        assert opt.param_groups[0]["lr"] == pytest.approx(ropt.param_groups[0]["lr"])
        assert opt.param_groups[0]["lr"] / cfg["m_width"] == pytest.approx(ropt.param_groups[1]["lr"])
    after = sdp.state_dict()
    for n in sd:
        assert torch.allclose(after[n], ref[n].data, rtol=1e-5, atol=1e-7), n
        assert not torch.equal(after[n], sd[n])
    # the bf16 compute copy (what the kernels read at world size 1) was written range by range
    for n, u, s in eng.named_views():
        assert torch.equal(u.views[n].float(), after[n].bfloat16().float()), n
    # several launches per block unit (alternating multipliers), learning rates lr and lr / m_width only
    assert len(calls) > 5 * len(eng.units) and {round(r, 6) for _, r in calls[-8:]} == {1.0, 0.25}


def test_mup_groups_refuse_what_the_reference_refuses():
    from dolomite_engine_b200.distributed import ShardedDataParallel
    from dolomite_engine_b200.model_wrapper import ModelWrapperForPretraining

    oargs = {"lr": 1e-3}
    normal = ModelWrapperForPretraining(pretrained_config=dict(KW, model_type="gpt_dolomite", init_method="normal"), micro_batch_size=1,
                                        sequence_length=16, device=torch.device("cpu"))
    with pytest.raises(AssertionError, match="init method"):
        get_optimizer("DolomiteFusedAdamW", oargs, ShardedDataParallel(normal, None), params_group_method="mup")
    mup = ShardedDataParallel(ModelWrapperForPretraining(pretrained_config=dict(KW, model_type="gpt_dolomite"), micro_batch_size=1,
                                                         sequence_length=16, device=torch.device("cpu")), None)
    with pytest.raises(NotImplementedError):
        get_optimizer("TorchAdamW", oargs, mup, params_group_method="mup")
    with pytest.raises(ValueError):
        get_optimizer("DolomiteFusedAdamW", oargs, mup, params_group_method="layerwise")
    assert get_optimizer("DolomiteFusedAdamW", oargs, mup).lr_scale_of is None  # default: one learning rate, one launch per shard
