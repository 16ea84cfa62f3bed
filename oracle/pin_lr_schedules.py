"""Pins optimization.get_scheduler (constant / cosine / exponential / linear / power) against the reference's
optimization/scheduler.py (runs only where /root/reference exists): the reference module is imported read-only under a
stubbed parent package and stepped next to ours; its learning-rate sequences are written to tests/golden/lr_schedules.json.

    python oracle/pin_lr_schedules.py

Test infrastructure only."""
import importlib.util
import json
import os
import sys
import types
from enum import Enum

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [("power", {"a": 4.6, "b": -0.51, "c": 2048.0}, 10, 0, None, 3e-2), ("power", {"a": 0.02, "b": -0.3, "c": 1.0}, 3, 0, None, 1e-2),
         ("cosine", {}, 5, 3, 20, 1e-3), ("cosine", {}, 0, 0, None, 1e-3), ("linear", {}, 4, 0, None, 1e-3),
         ("exponential", {}, 2, 2, 30, 1e-3), ("constant", {}, 3, 0, 0, 1e-3)]


def reference_module():
    class LRDecaySchedule(Enum):
        constant = "constant"
        cosine = "cosine"
        exponential = "exponential"
        linear = "linear"
        power = "power"

    for name in ("dolomite_engine", "dolomite_engine.optimization"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    en = types.ModuleType("dolomite_engine.enums")
    en.LRDecaySchedule = LRDecaySchedule
    sys.modules["dolomite_engine.enums"] = en
    spec = importlib.util.spec_from_file_location("dolomite_engine.optimization.scheduler",
                                                  "/root/reference/dolomite_engine/optimization/scheduler.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod, LRDecaySchedule


def main():
    from dolomite_engine_b200.optimization import get_scheduler

    ref_mod, Style = reference_module()
    out = []
    for style, extra, warm, const, decay, lr in CASES:
        o1 = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=lr)
        o2 = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=lr)
        ref = ref_mod.get_scheduler(o1, warm, const, decay, 40, Style(style), 0.1, extra)
        mine = get_scheduler(o2, warm, const, decay, 40, style, 0.1, extra)
        a, b = [], []
        for _ in range(45):
            a.append(ref.get_last_lr()[0])
            b.append(mine.get_last_lr()[0])
            o1.step(), o2.step(), ref.step(), mine.step()
        assert a == b, (style, max(abs(x - y) for x, y in zip(a, b)))
        out.append({"style": style, "extra": extra, "warmup": warm, "constant": const, "decay": decay, "lr": lr, "values": a})
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "lr_schedules.json"), "w"), indent=1)
    print("pinned", len(out), "schedules bit for bit")


if __name__ == "__main__":
    main()
