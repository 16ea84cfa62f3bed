"""Optimizer + LR schedule of the training step (reference: optimization/optimizer.py:55-164,
optimization/scheduler.py:50-219, defaults arguments.py:235-266).

`DolomiteFusedAdamW` runs torch.optim.AdamW arithmetic as ONE hand-written kernel per flat shard
(dolomite_b200_adamw_step): grad-clip coefficient read from a device scalar, fp32 master update and the bf16 copy
for the next all-gather in the same pass.  `TorchAdamW` is accepted for drop-in compatibility."""

from __future__ import annotations

import math

import torch
from torch.optim import Optimizer
from torch.optim.lr_scheduler import LambdaLR

from . import kernels as K


def lr_scale_segments(unit, scale_of) -> list[tuple[int, int, float]]:
    """[lo, hi) ranges of THIS RANK's shard of `unit` (offsets relative to the shard) with the learning-rate multiplier of the
    parameter they belong to, adjacent ranges of equal multiplier merged.  A parameter owns the alignment gap behind it (the gap
    holds zeros and zero gradients: any multiplier leaves it zero); a shard boundary may cut a parameter anywhere."""
    lo_s, hi_s = unit.rank * unit.shard_numel, (unit.rank + 1) * unit.shard_numel
    out: list[list] = []
    for i, s in enumerate(unit.specs):
        start = s.offset if i else 0
        end = unit.specs[i + 1].offset if i + 1 < len(unit.specs) else unit.padded
        lo, hi = max(start, lo_s), min(end, hi_s)
        if lo >= hi:
            continue
        sc = float(scale_of(s.name))
        if out and out[-1][2] == sc and out[-1][1] == lo - lo_s:
            out[-1][1] = hi - lo_s
        else:
            out.append([lo - lo_s, hi - lo_s, sc])
    if not out:  # a unit without parameters in this shard (padding only)
        out.append([0, unit.shard_numel, 1.0])
    return [tuple(x) for x in out]


def mup_lr_scale(config):
    """optimization/optimizer.py:86-126 (`params_group_method: mup`): every parameter of the Attention and MLP modules except
    biases trains with lr / m_width; embeddings, norms, biases and the head keep lr"""
    import re

    pat = re.compile(r"^transformer\.h\.\d+\.(attn|mlp)\..*(?<!bias)$")
    inv = 1.0 / float(config.m_width)
    return lambda name: inv if pat.match(name) else 1.0


class DolomiteFusedAdamW(Optimizer):
    def __init__(self, params, lr=1e-5, betas=(0.9, 0.95), eps=1e-10, weight_decay=0.1, model=None, lr_scale_of=None):
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.model = model  # ShardedDataParallel: provides the clip coefficient and the bf16 targets
        self._step = 0
        # parameter name -> learning-rate multiplier (muP parameter groups); None = one learning rate, one launch per shard
        self.lr_scale_of = lr_scale_of
        self._segments: dict[int, list[tuple[int, int, float]]] = {}

    def zero_grad(self, set_to_none: bool = False) -> None:
        """The reference's train_step calls `optimizer.zero_grad()` (train_utils.py:40).  The gradients this optimizer
        consumes live in the engine's flat buffers (full fp32 accumulation buffers + shard gradients aliased by
        `master.grad`), so clearing is the model's job; `set_to_none` would break the aliasing and is ignored."""
        self.model.zero_grad()

    @torch.no_grad()
    def step(self, closure=None):
        assert closure is None
        self._step += 1
        sdp = self.model
        engine = sdp.engine
        by_ptr = {u.master.data_ptr(): u for u in engine.units}
        clip = sdp.clip_coef
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                u = by_ptr[p.data_ptr()]
                # world_size == 1: write the bf16 compute copy in the same pass; sharded: into own slice of the
                # gather buffer is done by the all-gather prologue (cast), so skip here
                pb = u.compute if engine.world_size == 1 else None
                if self.lr_scale_of is None:
                    K.adamw_step(p.data, p.grad, st["exp_avg"], st["exp_avg_sq"], pb, group["lr"], b1, b2, group["eps"],
                                 group["weight_decay"], self._step, clip=clip)
                    continue
                # the reference puts the muP parameters into a second torch param group with lr / m_width; here the groups are
                # ranges of the flat shard: the same kernel runs once per range with the range's learning rate
                key = id(u)
                if key not in self._segments:
                    self._segments[key] = lr_scale_segments(u, self.lr_scale_of)
                for lo, hi, scale in self._segments[key]:
                    K.adamw_step(p.data[lo:hi], p.grad[lo:hi], st["exp_avg"][lo:hi], st["exp_avg_sq"][lo:hi],
                                 None if pb is None else pb[lo:hi], group["lr"] * scale, b1, b2, group["eps"],
                                 group["weight_decay"], self._step, clip=clip)
        sdp.clip_coef.fill_(1.0)
        sdp.notify_fused_update()


# the reference's registry (optimization/optimizer.py:55-84): every torch.optim class under its "Torch<Name>" key works on
# the flat fp32 shards unchanged; the Apex / DeepSpeed entries are other backends' fused kernels (no multi-backend
# dispatch here) -- DolomiteFusedAdamW is this engine's fused kernel and the default of the shipped configs
_OPTIMIZER_CLASSES = {
    f"Torch{name}": getattr(torch.optim, name)
    for name in ("Adadelta", "Adagrad", "Adam", "Adamax", "AdamW", "ASGD", "NAdam", "RAdam", "RMSprop", "Rprop", "SGD")
}
_OPTIMIZER_CLASSES["DolomiteFusedAdamW"] = DolomiteFusedAdamW


def get_optimizer(optimizer_class_name: str, optimizer_class_args: dict, model, params_group_method=None) -> Optimizer:
    """optimization/optimizer.py:129-164"""
    if optimizer_class_name not in _OPTIMIZER_CLASSES:
        raise ValueError(
            f"invalid class_name ({optimizer_class_name}) for optimizer; the B200 path provides {sorted(_OPTIMIZER_CLASSES)}"
        )
    args = dict(optimizer_class_args)
    if "betas" in args:
        args["betas"] = tuple(args["betas"])
    cls = _OPTIMIZER_CLASSES[optimizer_class_name]
    params = list(model.parameters())
    lr_scale_of = None
    method = None if params_group_method is None else str(getattr(params_group_method, "value", params_group_method))
    if method == "mup":
        # optimization/optimizer.py:89-99: same three conditions as the reference's asserts
        cfg = model.config
        assert cfg.model_type == "gpt_dolomite", "mup is not supported with this model architecture"
        assert cfg.init_method == "mup", "both init method for model and params group method for optimizer should be set to mup"
        if cls is not DolomiteFusedAdamW:
            raise NotImplementedError("params_group_method: mup needs class_name: DolomiteFusedAdamW here (the torch optimizers see "
                                      "one flat fp32 shard per unit, which cannot be split into torch param groups)")
        lr_scale_of = mup_lr_scale(cfg)
    elif method is not None:
        raise ValueError(f"unexpected params_group_method ({params_group_method})")
    if cls is DolomiteFusedAdamW:
        return cls(params, model=model, lr_scale_of=lr_scale_of, **args)
    return cls(params, **args)


def _linear(m, c, x):
    return m * x + c


def _cosine(a, b, t, x):
    return a * (1 + math.cos(math.pi * x / t)) / 2 + b


def _exponential(a, b, t, x):
    return a * math.exp(-x / t) + b


class _LRScheduler(LambdaLR):
    def __init__(self, optimizer, num_warmup_steps, num_constant_steps, num_decay_steps, num_training_steps,
                 lr_decay_factor, last_epoch=-1):
        self.lr_warmup_boundary = num_warmup_steps
        self.lr_constant_boundary = self.lr_warmup_boundary + num_constant_steps
        self.lr_decay_boundary = num_training_steps
        if num_decay_steps is not None:
            self.lr_decay_boundary = self.lr_constant_boundary + num_decay_steps
        self.lr_decay_factor = lr_decay_factor
        super().__init__(optimizer, lr_lambda=self._lr_lambda, last_epoch=last_epoch)

    def _decay(self, x, t):
        raise NotImplementedError

    def _lr_lambda(self, num_steps: int) -> float:
        if self.lr_warmup_boundary > 0 and num_steps <= self.lr_warmup_boundary:
            return _linear(m=1 / self.lr_warmup_boundary, c=0, x=num_steps)
        if num_steps <= self.lr_constant_boundary:
            return 1
        if num_steps <= self.lr_decay_boundary:
            return self._decay(num_steps - self.lr_constant_boundary, self.lr_decay_boundary - self.lr_constant_boundary)
        return self.lr_decay_factor


class ConstantScheduler(_LRScheduler):
    def __init__(self, optimizer, num_warmup_steps, num_constant_steps, num_decay_steps, num_training_steps, lr_decay_factor,
                 last_epoch=-1):
        assert num_decay_steps == 0, "num_decay_steps should be 0 for constant schedule"  # scheduler.py:61
        super().__init__(optimizer, num_warmup_steps, num_constant_steps, num_decay_steps, num_training_steps, lr_decay_factor,
                         last_epoch=last_epoch)

    def _lr_lambda(self, num_steps: int) -> float:
        if self.lr_warmup_boundary > 0 and num_steps <= self.lr_warmup_boundary:
            return _linear(m=1 / self.lr_warmup_boundary, c=0, x=num_steps)
        return 1


class CosineScheduler(_LRScheduler):
    def _decay(self, x, t):
        return _cosine(a=1 - self.lr_decay_factor, b=self.lr_decay_factor, t=t, x=x)


class ExponentialScheduler(_LRScheduler):
    """optimization/scheduler.py:101-116: a * exp(-x / t) + b with a, b chosen so that the multiplier is 1 at the start of the
    decay and `lr_decay_factor` after t steps; unlike cosine / linear it keeps following the exponential afterwards"""

    def _lr_lambda(self, num_steps: int) -> float:
        if self.lr_warmup_boundary > 0 and num_steps <= self.lr_warmup_boundary:
            return _linear(m=1 / self.lr_warmup_boundary, c=0, x=num_steps)
        if num_steps <= self.lr_constant_boundary:
            return 1
        f, e = self.lr_decay_factor, math.e
        return _exponential(a=(1 - f) * e / (e - 1), b=(f * e - 1) / (e - 1), t=self.lr_decay_boundary - self.lr_constant_boundary,
                            x=num_steps - self.lr_constant_boundary)


class LinearScheduler(_LRScheduler):
    def _decay(self, x, t):
        return _linear(m=(self.lr_decay_factor - 1) / t, c=1, x=x)


class PowerScheduler(_LRScheduler):
    """optimization/scheduler.py:137-181 (the Granite "power" schedule): after a linear warm-up the multiplier is
    min(1, (a / lr) * (c * step)^b) with `a, b, c` from `extra_lr_scheduler_args`; no constant phase, `lr_decay_factor` unused.
    The warm-up ramps to the value the power law has at its last warm-up step, so the two pieces meet."""

    def __init__(self, optimizer, num_warmup_steps, num_constant_steps, num_decay_steps, num_training_steps, lr_decay_factor,
                 a: float, b: float, c: float, last_epoch=-1):
        assert num_constant_steps == 0, "num_constant_steps should be 0 for power law scheduler"
        self.a, self.b, self.c = a, b, c
        self._lr0 = optimizer.param_groups[0]["lr"]  # the optimizer's configured learning rate
        self._warmup_peak = self._law(num_warmup_steps)
        super().__init__(optimizer, num_warmup_steps, num_constant_steps, num_decay_steps, num_training_steps, lr_decay_factor,
                         last_epoch=last_epoch)

    def _law(self, step: int) -> float:
        return min(1, (self.a / self._lr0) * (step * self.c) ** self.b)

    def _lr_lambda(self, num_steps: int) -> float:
        if self.lr_warmup_boundary > 0 and num_steps <= self.lr_warmup_boundary:
            return _linear(m=self._warmup_peak / self.lr_warmup_boundary, c=0, x=num_steps)
        return self._law(num_steps)


_LR_SCHEDULER_CLASSES = {"constant": ConstantScheduler, "cosine": CosineScheduler, "exponential": ExponentialScheduler,
                         "linear": LinearScheduler, "power": PowerScheduler}


def get_scheduler(optimizer, num_warmup_steps, num_constant_steps, num_decay_steps, num_training_steps, lr_decay_style,
                  lr_decay_factor, extra_lr_scheduler_args=None, last_epoch=-1) -> LambdaLR:
    """optimization/scheduler.py:193-219"""
    style = str(getattr(lr_decay_style, "value", lr_decay_style))
    if style not in _LR_SCHEDULER_CLASSES:
        raise ValueError(f"invalid lr_decay_style ({lr_decay_style})")
    return _LR_SCHEDULER_CLASSES[style](
        optimizer, num_warmup_steps=num_warmup_steps, num_constant_steps=num_constant_steps,
        num_decay_steps=num_decay_steps, num_training_steps=num_training_steps, lr_decay_factor=lr_decay_factor,
        **(extra_lr_scheduler_args or {}), last_epoch=last_epoch,
    )
