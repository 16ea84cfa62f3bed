"""`fsdp_algorithm: 2` checkpoint flavour of the reference (checkpointing.py:108-113, :196-211): the model and the
optimizer state are `torch.distributed.checkpoint` DIRECTORIES

    <save_path>/global_step<N>/model/        dcp.save({fqn: tensor})
    <save_path>/global_step<N>/optimizer/    dcp.save({"state": {fqn: {step, exp_avg, exp_avg_sq}}, "param_groups": [...]})

keyed by the reference's fully-qualified parameter names.  The reference saves DTensor shards; a DCP checkpoint is
layout independent (a reader asks for whatever slices its own sharding needs), so this writer stores every named tensor
whole, and spreads the units round-robin over the ranks (rank r writes the tensors of units r, r + world, ...): the flat
fp32 shards of a unit are all-gathered once, the owner cuts them into named tensors, and all ranks write their files in
parallel.  Loading asks DCP for full named tensors unit by unit (bounded host memory) and scatters them into the shards.
"""

from __future__ import annotations

import os

import torch
import torch.distributed as dist
import torch.distributed.checkpoint as dcp

from .checkpointing import _PREFIX, _engine, _gather_flat, _named, _rank_world, _scatter_named


def _owner(unit_index: int, world: int) -> int:
    return unit_index % world


def save_model(model, path: str) -> None:
    """collective: every rank calls it"""
    engine = _engine(model)
    rank, world = _rank_world()
    mine: dict[str, torch.Tensor] = {}
    for i, u in enumerate(engine.units):
        full = _gather_flat(engine, u, u.master.data)
        if _owner(i, world) == rank:
            mine.update(_named(u, full))
        del full
    dcp.save(mine, checkpoint_id=path)


def save_optimizer(model, optimizer, path: str) -> None:
    engine = _engine(model)
    rank, world = _rank_world()
    step = int(getattr(optimizer, "_step", 0))
    state: dict[str, dict] = {}
    for i, u in enumerate(engine.units):
        st = optimizer.state.get(u.master, {})
        if "step" in st:
            step = int(float(st["step"]))
        moments = {}
        for k in ("exp_avg", "exp_avg_sq"):
            shard = st[k] if k in st else torch.zeros_like(u.master.data)
            full = _gather_flat(engine, u, shard)
            if _owner(i, world) == rank:
                moments[k] = _named(u, full)
            del full
        if _owner(i, world) == rank:
            for s in u.specs:
                key = _PREFIX + s.name
                state[key] = {"step": torch.tensor(float(step)), "exp_avg": moments["exp_avg"][key],
                              "exp_avg_sq": moments["exp_avg_sq"][key]}
    out: dict = {"state": state}
    if rank == 0:
        groups = []
        for g in optimizer.param_groups:
            d = {k: (list(v) if isinstance(v, tuple) else v) for k, v in g.items() if k != "params"}
            d["params"] = [_PREFIX + s.name for u in engine.units for s in u.specs]
            groups.append(d)
        out["param_groups"] = groups
    dcp.save(out, checkpoint_id=path)


def _blank(unit) -> dict[str, torch.Tensor]:
    return {_PREFIX + s.name: torch.zeros(s.shape, dtype=torch.float32) for s in unit.specs}


def load_model(model, path: str) -> None:
    """collective; one unit at a time so that at most one unit's parameters sit in host memory"""
    engine = _engine(model)
    for u in engine.units:
        named = _blank(u)
        dcp.load(named, checkpoint_id=path)
        full = torch.zeros(u.padded, dtype=torch.float32)
        for s in u.specs:
            full[s.offset : s.offset + s.numel] = named[_PREFIX + s.name].reshape(-1)
        u.full_master_from(full)  # this rank's fp32 slice + the whole bf16 compute copy


def load_optimizer(model, optimizer, path: str) -> None:
    engine = _engine(model)
    step = 0
    for u in engine.units:
        want = {"state": {k: {"step": torch.zeros(()), "exp_avg": torch.zeros_like(v), "exp_avg_sq": torch.zeros_like(v)}
                          for k, v in _blank(u).items()}}
        dcp.load(want, checkpoint_id=path)
        st = optimizer.state[u.master]
        for k in ("exp_avg", "exp_avg_sq"):
            if k not in st:
                st[k] = torch.zeros_like(u.master.data)
            _scatter_named(u, {key: v[k] for key, v in want["state"].items()}, st[k], f"optimizer checkpoint ({k})")
        step = int(float(next(iter(want["state"].values()))["step"]))
    if hasattr(optimizer, "_step"):
        optimizer._step = step
    else:
        for u in engine.units:
            optimizer.state[u.master]["step"] = torch.tensor(float(step))
    stored = dcp.FileSystemReader(path).read_metadata().state_dict_metadata
    if not any(k.startswith("param_groups") for k in stored):
        return  # a checkpoint written without hyper-parameters keeps the YAML's values
    groups = {"param_groups": [dict(g, params=[]) for g in optimizer.param_groups]}
    dcp.load(groups, checkpoint_id=path)
    for g, saved in zip(optimizer.param_groups, groups["param_groups"]):
        for k, v in saved.items():
            if k != "params":
                g[k] = tuple(v) if k == "betas" else v


def is_dcp_checkpoint(load_path: str) -> bool:
    return os.path.isdir(os.path.join(load_path, "model")) and not os.path.isfile(os.path.join(load_path, "model.pt"))
