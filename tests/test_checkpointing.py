"""Reference checkpoint layout (checkpointing.py:50-263): file tree, key names, full-tensor
round trip through the flat shards -- single process and world_size 2 (gloo, CPU)."""
import json
import os
import sys
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(world: int, rank: int, seed: int):
    from dolomite_engine_b200.engine import DolomiteEngine
    from dolomite_engine_b200.hf_models import GPTDolomiteConfig

    cfg = GPTDolomiteConfig(n_embd=64, n_head=4, n_layer=2, n_inner=96, vocab_size=264, attention_head_type="gqa",
                            num_key_value_heads=2, add_bias=True, activation_function="swiglu",
                            position_embedding_type="rope", normalization_function="rmsnorm", resid_pdrop=0, embd_pdrop=0,
                            attn_pdrop=0)
    engine = DolomiteEngine(cfg, "cpu", world_size=world, rank=rank, seed=seed)
    if world > 1:
        engine.comm = types.SimpleNamespace(group=dist.group.WORLD)
    model = types.SimpleNamespace(engine=engine)
    opt = torch.optim.AdamW([u.master for u in engine.units], lr=1e-3, betas=(0.9, 0.95), eps=1e-10, weight_decay=0.1)
    for ui, u in enumerate(engine.units):  # make the moments non-trivial (each rank owns a different slice of them)
        full = torch.zeros(u.padded)  # the alignment padding never receives a gradient
        full[: u.numel] = torch.randn(u.numel, generator=torch.Generator().manual_seed(100 + seed + ui))
        u.master.grad = full[rank * u.shard_numel : (rank + 1) * u.shard_numel].clone()
    opt.step()
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0 / (1 + s))
    sched.step()
    return engine, model, opt, sched


def _args(path, load=False, fsdp_algorithm=1):
    ns = types.SimpleNamespace
    return ns(save_args=ns(save_path=path, save_optimizer=True), distributed_args=ns(fsdp_algorithm=fsdp_algorithm),
              load_args=ns(load_path=path, iteration=None, load_optimizer=True, load_lr_scheduler=True, load_rng_state=True,
                           load_dataloader_state=True, load_experiments_tracker_state=True, load_starting_iteration=True) if load else None,
              model_dump=lambda mode="json": {"note": "training config"})


def _roundtrip(path, world, rank, fsdp_algorithm=1):
    from dolomite_engine_b200 import checkpointing as C

    engine, model, opt, sched = _build(world, rank, seed=1)
    loader = types.SimpleNamespace(consumed_samples=48)
    want_model = C.model_state_dict(model)
    want_opt = C.optimizer_state_dict(model, opt)
    C.save_checkpoint(_args(path, fsdp_algorithm=fsdp_algorithm), model, opt, sched, loader, None, 7,
                      metadata={"consumed_samples": 48})
    # a differently initialised replica restores everything
    engine2, model2, opt2, sched2 = _build(world, rank, seed=2)
    assert not torch.equal(engine2.units[1].master.data, engine.units[1].master.data)
    it, meta, _ = C.load_checkpoint_for_training(_args(path, load=True), model2, opt2, sched2, None)
    assert it == 7 and meta == {"consumed_samples": 48}
    for u, u2 in zip(engine.units, engine2.units):
        assert torch.equal(u.master.data, u2.master.data)
        if world == 1:  # bf16 compute copy refreshed from the loaded masters
            assert torch.equal(u2.compute[: u2.shard_numel], u2.master.data.bfloat16())
        for k in ("exp_avg", "exp_avg_sq"):
            assert torch.equal(opt.state[u.master][k], opt2.state[u2.master][k])
        assert float(opt2.state[u2.master]["step"]) == 1.0
    assert sched2.state_dict()["last_epoch"] == sched.state_dict()["last_epoch"]
    got = C.model_state_dict(model2)
    assert got.keys() == want_model.keys() and all(torch.equal(got[k], want_model[k]) for k in got)
    return want_model, want_opt


def test_single_process_layout_and_names(tmp_path):
    path = str(tmp_path / "ckpt")
    sd, osd = _roundtrip(path, 1, 0)
    base = os.path.join(path, "global_step7")
    for rel in ("model.pt", "optimizer.pt", "lr_scheduler.pt", "rng_state/rng_state-0.pt", "dataloader/dataloader-0.pt",
                "metadata.json", "training_config.yml"):
        assert os.path.isfile(os.path.join(base, rel)), rel
    assert json.load(open(os.path.join(path, "latest_checkpointed_iteration.json"))) == {"latest_checkpointed_iteration": 7}
    on_disk = torch.load(os.path.join(base, "model.pt"))
    assert set(on_disk) == set(sd)
    # reference fully-qualified names under the wrapper's "model." prefix; tied head is not serialised
    for k in ("model.transformer.wte.weight", "model.transformer.h.0.ln_1.weight", "model.transformer.h.1.attn.c_attn.bias",
              "model.transformer.h.1.mlp.c_proj.weight", "model.transformer.ln_f.weight"):
        assert k in on_disk and on_disk[k].dtype == torch.float32
    assert "model.lm_head.weight" not in on_disk
    assert on_disk["model.transformer.h.0.attn.c_attn.weight"].shape == (64 + 2 * 2 * 16, 64)
    o = torch.load(os.path.join(base, "optimizer.pt"))
    assert set(o) == {"state", "param_groups"} and set(o["state"]) == set(sd)
    e = o["state"]["model.transformer.h.0.mlp.c_fc.weight"]
    assert set(e) == {"step", "exp_avg", "exp_avg_sq"} and e["exp_avg"].shape == (192, 64)
    assert o["param_groups"][0]["lr"] == pytest.approx(5e-4) and o["param_groups"][0]["params"][0].startswith("model.")
    assert torch.load(os.path.join(base, "dataloader", "dataloader-0.pt"), weights_only=False) == {"consumed_samples": 48}


def test_dropout_generator_state_travels_with_the_rng_state(tmp_path):
    """dropout masks are counter-based: (seed, passes so far) is the generator state; it is stored next to the torch / numpy /
    python RNG states (rng_state-<rank>.pt; the reference stores its Philox state there) so that a resumed run continues the mask
    sequence instead of repeating it"""
    from dolomite_engine_b200 import checkpointing as C

    path = str(tmp_path / "ckpt")
    engine, model, opt, sched = _build(1, 0, seed=1)
    engine.has_dropout, engine.dropout_seed, engine._dropout_passes = True, 4242, 17
    C.save_checkpoint(_args(path), model, opt, sched, None, None, 3, metadata={})
    engine2, model2, opt2, sched2 = _build(1, 0, seed=2)
    engine2.has_dropout = True
    C.load_checkpoint_for_training(_args(path, load=True), model2, opt2, sched2, None)
    assert (engine2.dropout_seed, engine2._dropout_passes) == (4242, 17)
    engine2.training = True
    engine2._begin_dropout_pass()
    assert engine2._dropout_now == 4242 + 17


def _worker(rank, world, port, path, q, fsdp_algorithm=1):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sd, _ = _roundtrip(path, world, rank, fsdp_algorithm)
        if fsdp_algorithm == 2 and rank == 0:
            _check_dcp_directory(os.path.join(path, "global_step7"), sd, world)
        # the sharded save must describe the same full tensors as an unsharded replica built from the same seed
        engine1, model1, _, _ = _build(1, 0, seed=1)
        from dolomite_engine_b200 import checkpointing as C

        ref = C.model_state_dict(model1)
        assert all(torch.equal(ref[k], sd[k]) for k in ref)
        q.put((rank, "ok"))
    except Exception:  # noqa
        import traceback

        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _check_dcp_directory(base, sd, world):
    """fsdp_algorithm 2: `model/` and `optimizer/` are torch.distributed.checkpoint directories keyed by the reference's
    names; any reader (here: a plain single-process dcp.load of full tensors) gets the same values back"""
    import torch.distributed.checkpoint as dcp
    from torch.distributed.checkpoint import FileSystemReader

    assert not os.path.exists(os.path.join(base, "model.pt")) and not os.path.exists(os.path.join(base, "optimizer.pt"))
    for d in ("model", "optimizer"):
        assert os.path.isfile(os.path.join(base, d, ".metadata"))
    files = [f for f in os.listdir(os.path.join(base, "model")) if f.endswith(".distcp")]
    assert len(files) == world  # every rank wrote its share of the units
    md = FileSystemReader(os.path.join(base, "model")).read_metadata()
    assert set(md.state_dict_metadata) == set(sd)
    omd = FileSystemReader(os.path.join(base, "optimizer")).read_metadata()
    k = "model.transformer.h.0.mlp.c_fc.weight"
    assert {f"state.{k}.exp_avg", f"state.{k}.exp_avg_sq", f"state.{k}.step"} <= set(omd.state_dict_metadata)
    assert any(key.startswith("param_groups") for key in omd.state_dict_metadata)


def test_dcp_single_process_roundtrip(tmp_path):
    path = str(tmp_path / "ckpt_dcp")
    sd, _ = _roundtrip(path, 1, 0, fsdp_algorithm=2)
    _check_dcp_directory(os.path.join(path, "global_step7"), sd, 1)
    # values: read every tensor back with plain dcp.load
    import torch.distributed.checkpoint as dcp

    got = {k: torch.zeros_like(v) for k, v in sd.items()}
    dcp.load(got, checkpoint_id=os.path.join(path, "global_step7", "model"))
    assert all(torch.equal(got[k], sd[k]) for k in sd)


@pytest.mark.parametrize("fsdp_algorithm", [1, 2])
def test_world_size_2_checkpoint_roundtrip(tmp_path, fsdp_algorithm):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() + 7 * fsdp_algorithm) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path / "ckpt2"), q, fsdp_algorithm)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status in results:
        assert status == "ok", f"rank {rank}: {status}"


@pytest.mark.parametrize("fsdp_algorithm", [1, 2])
def test_unshard_writes_a_pretrained_directory(tmp_path, fsdp_algorithm):
    """unshard.py of the reference: global_step<N>/model.pt (or the DCP directory model/) -> safetensors + config.json with
    the reference's names"""
    from dolomite_engine_b200 import checkpointing as C
    from dolomite_engine_b200.hf_models.config import GPTDolomiteConfig
    from dolomite_engine_b200.unshard import unshard
    from dolomite_engine_b200.utils.safetensors import SafeTensorsWeightsManager

    path = str(tmp_path / "ckpt")
    engine, model, opt, sched = _build(1, 0, seed=3)
    args = _args(path, fsdp_algorithm=fsdp_algorithm)
    args.model_dump = lambda mode="json": {"model_args": {"pretrained_config": engine.cfg.to_dict()}}
    C.save_checkpoint(args, model, opt, sched, None, None, 11)
    out = unshard(path, str(tmp_path / "hf"), None)
    sd = SafeTensorsWeightsManager(out).state_dict()
    want = C.model_state_dict(model)
    assert set(sd) == {k[len("model."):] for k in want}
    assert all(torch.equal(sd[k[len("model."):]], v) for k, v in want.items())
    cfg = GPTDolomiteConfig.from_pretrained(out)
    assert cfg.n_embd == 64 and cfg.num_key_value_heads == 2 and cfg.attention_head_type == "gqa"


@pytest.mark.parametrize("fsdp_algorithm", [1, 2])
def test_load_checkpoint_for_inference_rebuilds_the_model(tmp_path, fsdp_algorithm):
    """checkpointing.py:266-402: the model is rebuilt from the training_config.yml stored with the checkpoint (as a padded-batch
    finetuning wrapper, which generation needs) and receives the saved parameters -- from model.pt or the DCP directory"""
    from dolomite_engine_b200 import checkpointing as C
    from dolomite_engine_b200.arguments import get_args_from_dict
    from dolomite_engine_b200.model_wrapper import ModelWrapperForFinetuning, get_model

    d = str(tmp_path / "run")
    pc = dict(model_type="gpt_dolomite", n_embd=64, n_head=4, n_layer=2, n_inner=96, vocab_size=264, attention_head_type="gqa",
              num_key_value_heads=2, add_bias=True, activation_function="swiglu", position_embedding_type="rope",
              normalization_function="rmsnorm", resid_pdrop=0, embd_pdrop=0, attn_pdrop=0)
    targs = get_args_from_dict({
        "model_args": {"model_class": "AutoModelForCausalLM", "pretrained_config": pc, "use_padding_free_transformer": True},
        "tuning_args": {"tuning_method": "pretraining"},
        "datasets": [{"class_name": "SyntheticPackedDataset", "data_name": "s", "class_args": {"sequence_length": 32}}],
        "training_parameters": {"num_training_steps": 1, "micro_batch_size": 2, "eval_during_training": False},
        "save_args": {"save_path": d, "save_interval": 1}, "distributed_args": {"fsdp_algorithm": fsdp_algorithm},
        "random_args": {"seed": 5}, "mixed_precision_args": {"dtype": "bf16"}})
    trained = get_model(targs, device=torch.device("cpu"))
    C.save_checkpoint(targs, trained, None, None, None, None, 3)
    iargs = get_args_from_dict({"datasets": [{"class_name": "JSONLinesDataset", "data_name": "x", "class_args": {"data_path": d}}],
                                "load_args": {"load_path": d}, "generation_parameters": {"batch_size": 1, "max_new_tokens": 1},
                                "output_dir": d + "/o"}, "inference")
    model, args_ckpt, state = C.load_checkpoint_for_inference(iargs, device=torch.device("cpu"))
    assert isinstance(model, ModelWrapperForFinetuning) and not model.use_padding_free_transformer
    assert args_ckpt.save_args.save_path == d and (state is None) == (fsdp_algorithm == 2)
    for u, v in zip(trained.model.engine.units, model.model.engine.units):
        assert torch.equal(u.master.data, v.master.data) and torch.equal(u.compute, v.compute)


def test_resume_learning_rate_when_the_scheduler_is_not_loaded(tmp_path):
    """checkpointing.py:232-238, :419-445: with load_lr_scheduler false (and resume_learning_rate true) the YAML's schedule is
    rebuilt at the loaded iteration on top of the learning rates stored in the optimizer checkpoint"""
    from dolomite_engine_b200 import checkpointing as C
    from dolomite_engine_b200.optimization import get_scheduler

    path = str(tmp_path / "ckpt")
    engine, model, opt, sched = _build(1, 0, seed=1)  # lr 1e-3 scheduled by 1 / (1 + s): after one step lr = 5e-4
    C.save_checkpoint(_args(path), model, opt, sched, None, None, 7)
    engine2, model2, opt2, _ = _build(1, 0, seed=2)
    ns = types.SimpleNamespace
    args = _args(path, load=True)
    args.load_args.load_lr_scheduler = False
    args.load_args.resume_learning_rate = True
    args.lr_scheduler_args = ns(num_warmup_steps=2, num_constant_steps=0, num_decay_steps=None, lr_decay_style="cosine",
                                lr_decay_factor=0.1, extra_lr_scheduler_args={})
    args.training_parameters = ns(num_training_steps=20)
    live = get_scheduler(opt2, 2, 0, None, 20, "cosine", 0.1)
    it, _, _ = C.load_checkpoint_for_training(args, model2, opt2, live, None)
    assert it == 7 and live.last_epoch == 7
    assert live.base_lrs == [pytest.approx(5e-4)]  # the loaded optimizer's lr became the base rate of the new phase
    # the same schedule started from scratch with base lr 5e-4 and stepped 7 times yields the same multiplier
    probe = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=5e-4)
    ref = get_scheduler(probe, 2, 0, None, 20, "cosine", 0.1)
    for _ in range(7):
        probe.step()
        ref.step()
    live.step()
    ref.step()
    assert live.get_last_lr() == pytest.approx(ref.get_last_lr())
    assert "initial_lr" in opt2.param_groups[0]  # torch's LambdaLR set it; the temporary override was undone
    # resume_learning_rate false: the live scheduler is left untouched
    engine3, model3, opt3, _ = _build(1, 0, seed=3)
    args.load_args.resume_learning_rate = False
    untouched = get_scheduler(opt3, 2, 0, None, 20, "cosine", 0.1)
    C.load_checkpoint_for_training(args, model3, opt3, untouched, None)
    assert untouched.last_epoch == 0


def test_wrapper_save_pretrained_from_a_full_state_dict(tmp_path):
    """model_wrapper/base.py:138-149: `save_pretrained(path, state_dict)` strips the wrapper prefix and writes safetensors +
    config.json that `from_pretrained`-style readers understand"""
    from dolomite_engine_b200 import checkpointing as C
    from dolomite_engine_b200.arguments import get_args_from_dict
    from dolomite_engine_b200.hf_models.config import CommonConfig
    from dolomite_engine_b200.model_wrapper import get_model
    from dolomite_engine_b200.utils.safetensors import SafeTensorsWeightsManager

    pc = dict(model_type="gpt_dolomite", n_embd=64, n_head=4, n_layer=1, n_inner=96, vocab_size=264, attention_head_type="mha",
              add_bias=False, activation_function="swiglu", position_embedding_type="rope", normalization_function="rmsnorm",
              resid_pdrop=0, embd_pdrop=0, attn_pdrop=0)
    args = get_args_from_dict({
        "model_args": {"model_class": "AutoModelForCausalLM", "pretrained_config": pc, "use_padding_free_transformer": True},
        "tuning_args": {"tuning_method": "full_finetuning"},
        "datasets": [{"class_name": "JSONLinesDataset", "data_name": "s", "class_args": {"data_path": "x"}}],
        "training_parameters": {"num_training_steps": 1, "micro_batch_size": 2, "eval_during_training": False},
        "save_args": {"save_path": str(tmp_path), "save_interval": 1}, "mixed_precision_args": {"dtype": "bf16"}})
    w = get_model(args, device=torch.device("cpu"))
    sd = C.model_state_dict(w)
    w.save_pretrained(str(tmp_path / "a"), state_dict=dict(sd))
    got = SafeTensorsWeightsManager(str(tmp_path / "a")).state_dict()
    assert set(got) == {k[len("model."):] for k in sd} and all(torch.equal(got[k[len("model."):]], v) for k, v in sd.items())
    assert CommonConfig.from_pretrained(str(tmp_path / "a")).n_inner == 96
    with pytest.raises(AssertionError):
        w.save_pretrained(str(tmp_path / "b"), state_dict={"transformer.wte.weight": torch.zeros(1)})
