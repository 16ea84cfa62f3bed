"""GPU bring-up probe: runs each kernel check in its own subprocess (a trap/hang in one cannot take the
rest down) and writes one JSON line per case to gpurun_out/probe.jsonl.

    python tools/gpu_probe.py                 # all cases
    python tools/gpu_probe.py --only gemm     # cases whose name contains 'gemm'
    python tools/gpu_probe.py --case NAME     # run one case in-process (used by the driver mode)

References are computed with torch on the GPU (test infrastructure only).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def _t():
    import torch

    return torch


def _err(a, b):
    a = a.float()
    b = b.float()
    d = (a - b).abs()
    return {
        "max_abs": d.max().item(),
        "rel_l2": (d.norm() / (b.norm() + 1e-30)).item(),
        "ref_absmax": b.abs().max().item(),
    }


def _time(fn, iters=20, warmup=3):
    torch = _t()
    iters = int(os.environ.get("PROBE_ITERS", iters))      # ncu runs: one launch per kernel is enough
    warmup = int(os.environ.get("PROBE_WARMUP", warmup))
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


# ---------------------------------------------------------------------------------------------
def _gemm_case(M, N, K, a_mn, b_mn, flags=0, out_f32=False, with_c=False, bias=False, alpha=1.0, beta=1.0, bench=False):
    torch = _t()
    from dolomite_engine_b200 import kernels as k

    g = torch.Generator(device="cuda").manual_seed(1)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.5).bfloat16()
    a_in = A.t().contiguous() if a_mn else A
    b_in = B.t().contiguous() if b_mn else B
    dt = torch.float32 if out_f32 else torch.bfloat16
    C = (torch.randn(M, N, device="cuda", generator=g)).to(dt) if with_c else None
    bv = (torch.randn(N, device="cuda", generator=g)).bfloat16() if bias else None
    ref = A.float() @ B.float().t()
    if bias:
        ref = ref + bv.float()
    ref = alpha * ref  # D = alpha * (A.B^T + bias) + beta * C
    if with_c:
        ref = ref + beta * C.float()
    out = k.gemm(a_in, b_in, a_mn=a_mn, b_mn=b_mn, out_dtype=dt, c=C, alpha=alpha, beta=beta, bias=bv, flags=flags)
    torch.cuda.synchronize()
    res = _err(out, ref)
    res["shape"] = [M, N, K]
    if bench:
        ms = _time(lambda: k.gemm(a_in, b_in, a_mn=a_mn, b_mn=b_mn, out=out, c=C, alpha=alpha, beta=beta, bias=bv, flags=flags))
        res["ms"] = ms
        res["tflops"] = 2.0 * M * N * K / ms / 1e9
        ms_ref = _time(lambda: torch.matmul(A, B.t()))
        res["cublas_ms"] = ms_ref
        res["cublas_tflops"] = 2.0 * M * N * K / ms_ref / 1e9
    tol = 2e-2 if not out_f32 else 1e-3
    res["ok"] = bool(res["rel_l2"] < tol)
    return res


@case
def gemm_nt_small():
    return _gemm_case(128, 256, 64, False, False)


@case
def gemm_nt_k256():
    return _gemm_case(128, 256, 256, False, False)


@case
def gemm_nt_multi_tile():
    return _gemm_case(512, 1024, 512, False, False)


@case
def gemm_nt_ragged():
    return _gemm_case(300, 520, 328, False, False)


@case
def gemm_nt_tma_store():
    return _gemm_case(512, 1024, 512, False, False, flags=1)


@case
def gemm_nt_tma_store_ragged():
    return _gemm_case(300, 520, 328, False, False, flags=1, bias=True)


@case
def gemm_nt_bias_c():
    return _gemm_case(384, 512, 256, False, False, with_c=True, bias=True, alpha=0.5, beta=1.0)


@case
def gemm_nn_bmn():
    return _gemm_case(512, 768, 512, False, True)


@case
def gemm_tn_amn():
    return _gemm_case(512, 768, 512, True, False)


@case
def gemm_tt_wgrad_f32_accum():
    return _gemm_case(640, 512, 1024, True, True, out_f32=True, with_c=True, beta=1.0)


@case
def gemm_many_tiles_persistent():
    # > 148 tiles so every CTA loops; exercises accumulator double buffering and phase flips
    return _gemm_case(4096, 4096, 1024, False, False)


@case
def gemm_bench_fwd_qkv():
    return _gemm_case(8192, 7680, 2560, False, False, bench=True)


@case
def gemm_bench_fwd_qkv_tma_store():
    return _gemm_case(8192, 7680, 2560, False, False, flags=1, bench=True)


@case
def gemm_bench_fc():
    return _gemm_case(8192, 20480, 2560, False, False, flags=1, bench=True)


@case
def gemm_bench_proj():
    return _gemm_case(8192, 2560, 10240, False, False, flags=1, bench=True)


@case
def gemm_bench_dgrad():
    return _gemm_case(8192, 2560, 7680, False, True, flags=1, bench=True)


@case
def gemm_bench_wgrad():
    return _gemm_case(20480, 2560, 8192, True, True, out_f32=True, with_c=True, bench=True)


@case
def gemm_pair_correctness():
    """CTA-pair (cta_group::2) kernel: every operand layout / epilogue, ragged sizes, many tiles"""
    out = {}
    ok = True
    cases = {
        "nt_256": (256, 256, 64, False, False, dict(flags=4)),
        "nt_tma": (512, 1024, 512, False, False, dict(flags=5)),
        "nt_ragged_bias": (300, 520, 328, False, False, dict(flags=5, bias=True)),
        "nt_bias_c": (384, 512, 256, False, False, dict(flags=4, with_c=True, bias=True, alpha=0.5)),
        "nn_bmn": (512, 768, 512, False, True, dict(flags=5)),
        "tn_amn": (512, 768, 512, True, False, dict(flags=5)),
        "tt_wgrad": (640, 512, 1024, True, True, dict(flags=4, out_f32=True, with_c=True)),
        "many_tiles": (4096, 4096, 1024, False, False, dict(flags=5)),
        "odd_super_tiles": (128 * 7, 256 * 3, 192, False, False, dict(flags=5)),
    }
    for name, (M, N, Kd, a_mn, b_mn, kw) in cases.items():
        r = _gemm_case(M, N, Kd, a_mn, b_mn, **kw)
        out[name] = r["rel_l2"]
        ok = ok and r["ok"]
    out["ok"] = ok
    return out


@case
def gemm_pair_bench():
    out = {}
    for name, (M, N, Kd, a_mn, b_mn, kw) in {
        "fwd_qkv": (8192, 7680, 2560, False, False, dict(flags=5)),
        "fwd_fc": (8192, 20480, 2560, False, False, dict(flags=5)),
        "fwd_proj": (8192, 2560, 10240, False, False, dict(flags=5)),
        "dgrad": (8192, 2560, 7680, False, True, dict(flags=5)),
        "wgrad_fc": (20480, 2560, 8192, True, True, dict(flags=4, out_f32=True, with_c=True)),
        "wgrad_qkv": (7680, 2560, 8192, True, True, dict(flags=4, out_f32=True, with_c=True)),
    }.items():
        r = _gemm_case(M, N, Kd, a_mn, b_mn, bench=True, **kw)
        base = _gemm_case(M, N, Kd, a_mn, b_mn, bench=True, **{**kw, "flags": kw["flags"] & ~4 | 8})
        out[name] = {"pair_tflops": round(r["tflops"]), "single_tflops": round(base["tflops"]), "cublas_tflops": round(r["cublas_tflops"]),
                     "rel_l2": r["rel_l2"]}
    # the long-contraction shapes again without the L2 eviction hints (A/B inside one process)
    from dolomite_engine_b200 import kernels as k_

    k_.set_option("gemm_l2_hints", 0)
    for name, (M, N, Kd, a_mn, b_mn, kw) in {
        "dgrad_fc_T24576": (24576, 2560, 20480, False, True, dict(flags=5)),
        "wgrad_fc_T24576": (20480, 2560, 24576, True, True, dict(flags=4, out_f32=True, with_c=True)),
    }.items():
        base = _gemm_case(M, N, Kd, a_mn, b_mn, bench=True, **kw)
        k_.set_option("gemm_l2_hints", 1)
        r = _gemm_case(M, N, Kd, a_mn, b_mn, bench=True, **kw)
        k_.set_option("gemm_l2_hints", 0)
        out[name] = {"hints_tflops": round(r["tflops"]), "no_hints_tflops": round(base["tflops"]), "cublas_tflops": round(r["cublas_tflops"]),
                     "rel_l2": r["rel_l2"]}
    k_.set_option("gemm_l2_hints", 1)
    out["ok"] = all(v["rel_l2"] < 2e-2 for v in out.values())
    return out


@case
def gemm_dynamic():
    """cluster-launch-control tile scheduling (gemm_dynamic = 1) against the static persistent schedule: exactness on every
    operand layout / epilogue, interleaved timing on the C2 shapes at T = 24576, and both next to a kernel that HOLDS 8 SMs
    (stand-in for a concurrent NCCL collective): static grid over all SMs, static grid with the 8-SM margin, dynamic"""
    import ctypes

    torch = _t()
    from dolomite_engine_b200 import _lib
    from dolomite_engine_b200 import kernels as k

    res = {}
    ok = True
    default_dynamic = k.get_option("gemm_dynamic")
    shapes = {
        "nt_256": (256, 256, 64, False, False, dict(flags=4)),
        "nt_tma": (512, 1024, 512, False, False, dict(flags=5)),
        "nt_ragged_bias": (300, 520, 328, False, False, dict(flags=5, bias=True)),
        "nt_bias_c": (384, 512, 256, False, False, dict(flags=4, with_c=True, bias=True, alpha=0.5)),
        "nn_bmn": (512, 768, 512, False, True, dict(flags=5)),
        "tn_amn": (512, 768, 512, True, False, dict(flags=5)),
        "tt_wgrad": (640, 512, 1024, True, True, dict(flags=4, out_f32=True, with_c=True)),
        "many_tiles": (8192, 8192, 512, False, False, dict(flags=5)),
        "odd_super_tiles": (128 * 7, 256 * 3, 192, False, False, dict(flags=5)),
        "single_many_tiles": (4096 + 64, 4096, 512, False, False, dict(flags=8)),
        "single_wgrad": (640, 512, 1024, True, True, dict(flags=8, out_f32=True, with_c=True)),
        "single_row_tile": (100, 2560, 2560, False, False, dict(flags=0)),
    }
    try:
        k.set_option("gemm_dynamic", 1)
        for name, (M, N, Kd, a_mn, b_mn, kw) in shapes.items():
            r = _gemm_case(M, N, Kd, a_mn, b_mn, **kw)
            res["dyn_" + name] = r["rel_l2"]
            ok = ok and r["ok"]
        # bit-identical to the static schedule (same tiles, same arithmetic)
        g = torch.Generator(device="cuda").manual_seed(3)
        A = (torch.randn(4096, 2560, device="cuda", generator=g) * 0.5).bfloat16()
        B = (torch.randn(7680, 2560, device="cuda", generator=g) * 0.5).bfloat16()
        d_dyn = k.gemm(A, B, flags=5)
        k.set_option("gemm_dynamic", 0)
        d_sta = k.gemm(A, B, flags=5)
        res["bit_identical"] = bool(torch.equal(d_dyn, d_sta))
        ok = ok and res["bit_identical"]

        # ---- timing, interleaved ----
        T, H, F = 24576, 2560, 10240
        x = (torch.randn(T, H, device="cuda", generator=g) * 0.1).bfloat16()
        w_fc = (torch.randn(2 * F, H, device="cuda", generator=g) * 0.02).bfloat16()
        y_fc = torch.empty(T, 2 * F, device="cuda", dtype=torch.bfloat16)
        w_attn = (torch.randn(3 * H, H, device="cuda", generator=g) * 0.02).bfloat16()
        y_attn = torch.empty(T, 3 * H, device="cuda", dtype=torch.bfloat16)
        dx = torch.empty(T, H, device="cuda", dtype=torch.bfloat16)
        a_proj = (torch.randn(T, F, device="cuda", generator=g) * 0.1).bfloat16()
        w_proj = (torch.randn(H, F, device="cuda", generator=g) * 0.02).bfloat16()
        probs = []
        for M, N in ((H, F), (2 * F, H), (H, H), (3 * H, H)):
            dy = (torch.randn(T, M, device="cuda", generator=g) * 0.1).bfloat16()
            xx = (torch.randn(T, N, device="cuda", generator=g) * 0.1).bfloat16()
            probs.append((dy, xx, torch.zeros(M, N, device="cuda"), 1.0, True))
        work = {
            "fwd_fc": (lambda: k.gemm(x, w_fc, out=y_fc, flags=5), 2.0 * T * 2 * F * H),
            "fwd_qkv": (lambda: k.gemm(x, w_attn, out=y_attn, flags=5), 2.0 * T * 3 * H * H),
            "fwd_proj": (lambda: k.gemm(a_proj, w_proj, out=dx, flags=5), 2.0 * T * H * F),
            "dgrad_fc": (lambda: k.gemm(y_fc, w_fc, b_mn=True, out=dx, flags=5), 2.0 * T * 2 * F * H),
            "wgrad_block": (lambda: k.gemm_wgrad_multi(probs), sum(2.0 * T * p[2].shape[0] * p[2].shape[1] for p in probs)),
        }
        for name, (fn, flops) in work.items():
            ts = {0: [], 1: []}
            for _ in range(3):
                for dyn in (0, 1):
                    k.set_option("gemm_dynamic", dyn)
                    ts[dyn].append(_time(fn, iters=10, warmup=2))
            res[name] = {"static_tflops": round(flops / min(ts[0]) / 1e9), "dynamic_tflops": round(flops / min(ts[1]) / 1e9),
                         "static_ms": ts[0], "dynamic_ms": ts[1]}

        # ---- next to a kernel that holds 8 SMs ----
        lib = _lib.load()
        hold = lib.dolomite_b200_debug_hold_sms
        hold.argtypes = [ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p]
        hold.restype = ctypes.c_int
        side = torch.cuda.Stream()
        fn, flops = work["fwd_fc"]

        def timed(hold_ms, n=6):
            torch.cuda.synchronize()
            if hold_ms > 0:
                rc = hold(8, int(hold_ms * 1.9e6), side.cuda_stream)
                assert rc == 0
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n

        for label, dyn, margin in (("static_all_sms", 0, 0), ("static_margin8", 0, 8), ("dynamic", 1, 0)):
            k.set_option("gemm_dynamic", dyn)
            k.set_option("gemm_sm_margin", margin)
            timed(0)
            res["hold_" + label] = {"alone_ms": min(timed(0) for _ in range(2)),
                                    "held_whole_time_ms": min(timed(40) for _ in range(2)),
                                    "held_first_third_ms": min(timed(3.5) for _ in range(2))}
    finally:
        k.set_option("gemm_dynamic", default_dynamic)
        k.set_option("gemm_sm_margin", 0)
    res["ok"] = ok
    return res


@case
def overlap_wgrad_elementwise():
    """Can the HBM-bound kernels of a block's backward hide behind the block's weight-gradient GEMMs?  A GEMM CTA leaves ~1.5 KB
    of shared memory, 39 K registers and 1800 thread slots of its SM free, so an elementwise CTA can be CO-RESIDENT with it.
    Serial (one stream) against concurrent (weight gradients on a second stream) time of: the four weight gradients of a C2
    block (one launch) + the block's elementwise backward kernels (SwiGLU bwd, 2 x RMSNorm bwd, RoPE bwd) at T = 24576."""
    torch = _t()
    from dolomite_engine_b200 import kernels as k

    T, H, F, nh, hd = 24576, 2560, 10240, 32, 80
    g = torch.Generator(device="cuda").manual_seed(0)
    probs = []
    for M, N in ((H, F), (2 * F, H), (H, H), (3 * H, H)):
        dy = (torch.randn(T, M, device="cuda", generator=g) * 0.1).bfloat16()
        xx = (torch.randn(T, N, device="cuda", generator=g) * 0.1).bfloat16()
        probs.append((dy, xx, torch.zeros(M, N, device="cuda"), 1.0, True))
    fc = (torch.randn(T, 2 * F, device="cuda", generator=g)).bfloat16()
    d_act = (torch.randn(T, F, device="cuda", generator=g)).bfloat16()
    x = (torch.randn(T, H, device="cuda", generator=g)).bfloat16()
    dy = (torch.randn(T, H, device="cuda", generator=g)).bfloat16()
    w = torch.ones(H, device="cuda").bfloat16()
    rstd = torch.ones(T, device="cuda")
    dw = torch.zeros(H, device="cuda")
    qkv = (torch.randn(T, 3 * H, device="cuda", generator=g)).bfloat16()
    cos = torch.ones(4096, hd, device="cuda").bfloat16()
    sin = torch.zeros(4096, hd, device="cuda").bfloat16()
    pos = (torch.arange(T, device="cuda") % 4096).long()
    d_fc = torch.empty_like(fc)
    dx = torch.empty_like(x)

    def elementwise():
        k.swiglu_bwd(d_act, fc, out=d_fc)
        k.rmsnorm_bwd(dy, x, w, rstd, dw, dx_add=dy, out=dx)
        k.rope_qk_inplace(qkv, nh, 1, hd, cos, sin, pos, inverse=True)
        k.rmsnorm_bwd(dy, x, w, rstd, dw, dx_add=dy, out=dx)

    def wgrad():
        k.gemm_wgrad_multi(probs)

    side = torch.cuda.Stream()
    main = torch.cuda.current_stream()

    def serial():
        wgrad()
        elementwise()

    def concurrent():
        side.wait_stream(main)
        with torch.cuda.stream(side):
            wgrad()
        elementwise()
        main.wait_stream(side)

    res = {}
    for name, fn in (("wgrad_only", wgrad), ("elementwise_only", elementwise), ("serial", serial), ("concurrent", concurrent),
                     ("serial_again", serial), ("concurrent_again", concurrent)):
        res[name + "_ms"] = _time(fn, iters=10, warmup=3)
    res["hidden_fraction_of_elementwise"] = (res["serial_ms"] - res["concurrent_ms"]) / res["elementwise_only_ms"]
    res["ok"] = True
    return res


@case
def gemm_bench_wgrad_splitk():
    """weight-gradient shapes of one C2 block: accumulate-in-place epilogue vs split-K atomic epilogue"""
    torch = _t()
    from dolomite_engine_b200 import kernels as k

    res = {}
    ok = True
    T = 8192
    for (n_out, k_in) in [(2560, 2560), (7680, 2560), (20480, 2560), (2560, 10240)]:
        dy = (torch.randn(T, n_out, device="cuda") * 0.1).bfloat16()
        x = (torch.randn(T, k_in, device="cuda") * 0.1).bfloat16()
        ref = dy.float().t() @ x.float()
        entry = {}
        for name, flag in (("rmw", 0), ("splitk", 2)):
            d = torch.zeros(n_out, k_in, device="cuda")
            k.gemm(dy, x, a_mn=True, b_mn=True, out=d, c=d, beta=1.0, flags=flag)
            torch.cuda.synchronize()
            e = _err(d, ref)
            ms = _time(lambda: k.gemm(dy, x, a_mn=True, b_mn=True, out=d, c=d, beta=1.0, flags=flag), iters=10)
            entry[name] = {"rel_l2": e["rel_l2"], "ms": ms, "tflops": 2.0 * T * n_out * k_in / ms / 1e9}
            ok = ok and e["rel_l2"] < 1e-4
        ms_ref = _time(lambda: torch.matmul(dy.t(), x), iters=10)
        entry["cublas_tflops"] = 2.0 * T * n_out * k_in / ms_ref / 1e9
        res[f"{n_out}x{k_in}"] = entry
    res["ok"] = bool(ok)
    return res


# ---------------------------------------------------------------------------------------------
@case
def rmsnorm():
    torch = _t()
    from dolomite_engine_b200 import kernels as k

    out = {}
    ok = True
    for T, H in [(256, 256), (1000, 2560), (64, 4096), (33, 8192)]:
        g = torch.Generator(device="cuda").manual_seed(2)
        x = torch.randn(T, H, device="cuda", generator=g).bfloat16()
        w = (1 + 0.1 * torch.randn(H, device="cuda", generator=g)).bfloat16()
        dy = torch.randn(T, H, device="cuda", generator=g).bfloat16()
        dres = torch.randn(T, H, device="cuda", generator=g).bfloat16()
        eps = 1e-5
        y, rstd = k.rmsnorm_fwd(x, w, eps)
        xf = x.float().requires_grad_(True)
        wf = w.float().requires_grad_(True)
        var = xf.pow(2).mean(-1, keepdim=True)
        xn = xf * torch.rsqrt(var + eps)
        yref = (w * xn.detach().to(torch.bfloat16)).float()
        e1 = _err(y, yref)
        yr2 = wf * xn
        yr2.backward(dy.float())
        dw = torch.zeros(H, device="cuda", dtype=torch.float32)
        dx = k.rmsnorm_bwd(dy, x, w, rstd, dw, dx_add=dres)
        e2 = _err(dx, xf.grad + dres.float())
        e3 = _err(dw, wf.grad)
        out[f"{T}x{H}"] = {"fwd": e1, "dx": e2, "dw": e3}
        ok = ok and e1["rel_l2"] < 5e-3 and e2["rel_l2"] < 1e-2 and e3["rel_l2"] < 1e-2
    # bandwidth
    T, H = 16384, 2560
    x = torch.randn(T, H, device="cuda").bfloat16()
    w = torch.ones(H, device="cuda").bfloat16()
    y = torch.empty_like(x)
    ms = _time(lambda: k.rmsnorm_fwd(x, w, 1e-5, out=y))
    out["fwd_GBps"] = 4.0 * T * H / ms / 1e6
    out["ok"] = bool(ok)
    return out


@case
def rope():
    torch = _t()
    from dolomite_engine_b200 import kernels as k

    out = {}
    ok = True
    for (T, ng, g, hd) in [(128, 4, 1, 64), (200, 32, 1, 80), (77, 8, 4, 128), (50, 1, 4, 32)]:
        gen = torch.Generator(device="cuda").manual_seed(3)
        width = ng * (g + 2) * hd
        qkv = torch.randn(T, width, device="cuda", generator=gen).bfloat16()
        npos = 512
        inv_freq = 1.0 / (10000 ** (torch.arange(0, hd, 2, dtype=torch.float32, device="cuda") / hd))
        t = torch.arange(npos, dtype=torch.float32, device="cuda")
        freqs = torch.outer(t, inv_freq)
        emb = torch.cat((freqs, freqs), dim=-1)
        cos = emb.cos().to(torch.bfloat16)
        sin = emb.sin().to(torch.bfloat16)
        pos = torch.randint(0, npos, (T,), device="cuda", generator=gen)
        v = qkv.view(T, ng, g + 2, hd).clone()
        c = cos[pos].unsqueeze(1).unsqueeze(1)
        s = sin[pos].unsqueeze(1).unsqueeze(1)
        rot = v[:, :, : g + 1]
        x1, x2 = torch.chunk(rot, 2, dim=-1)
        ref_rot = (rot * c) + (torch.cat((-x2, x1), dim=-1) * s)
        ref = v.clone()
        ref[:, :, : g + 1] = ref_rot
        got = k.rope_qk_inplace(qkv.clone(), ng, g, hd, cos, sin, pos)
        e = _err(got.view(T, ng, g + 2, hd), ref)
        # inverse(forward(x)) ~= x on rotated slots (orthogonality, up to bf16 rounding)
        back = k.rope_qk_inplace(got.clone(), ng, g, hd, cos, sin, pos.int(), inverse=True)
        e2 = _err(back, qkv)
        out[f"{T}_{ng}_{g}_{hd}"] = {"fwd": e, "roundtrip": e2}
        ok = ok and e["max_abs"] <= 4e-2 and e["rel_l2"] < 3e-3 and e2["rel_l2"] < 2e-2
    out["ok"] = bool(ok)
    return out


@case
def swiglu():
    torch = _t()
    from dolomite_engine_b200 import kernels as k

    T, F = 300, 1024
    x = torch.randn(T, 2 * F, device="cuda").bfloat16()
    dy = torch.randn(T, F, device="cuda").bfloat16()
    y = k.swiglu_fwd(x)
    xf = x.float().requires_grad_(True)
    u, g = xf.chunk(2, dim=-1)
    yr = u * torch.nn.functional.silu(g)
    yr.backward(dy.float())
    dx = k.swiglu_bwd(dy, x)
    e1, e2 = _err(y, yr), _err(dx, xf.grad)
    T, F = 8192, 10240
    xb = torch.randn(T, 2 * F, device="cuda").bfloat16()
    yb = torch.empty(T, F, device="cuda", dtype=torch.bfloat16)
    ms = _time(lambda: k.swiglu_fwd(xb, out=yb))
    return {"fwd": e1, "bwd": e2, "fwd_GBps": 6.0 * T * F / ms / 1e6, "ok": bool(e1["rel_l2"] < 5e-3 and e2["rel_l2"] < 5e-3)}


@case
def embedding_ce():
    torch = _t()
    from dolomite_engine_b200 import kernels as k

    V, H, T = 2048, 256, 500
    wte = torch.randn(V, H, device="cuda").bfloat16()
    ids = torch.randint(0, V, (T,), device="cuda")
    y = k.embedding_fwd(ids, wte, 1.0)
    e1 = _err(y, wte[ids])
    y2 = k.embedding_fwd(ids, wte, 12.0)
    e1b = _err(y2, (wte[ids].float() * 12.0))
    dout = torch.randn(T, H, device="cuda").bfloat16()
    dw = torch.zeros(V, H, device="cuda")
    k.embedding_bwd(ids, dout, dw, 1.0)
    ref = torch.zeros(V, H, device="cuda").index_add_(0, ids, dout.float())
    e2 = _err(dw, ref)
    # cross entropy
    res = {"emb_fwd": e1, "emb_fwd_scaled": e1b, "emb_bwd": e2}
    ok = e1["max_abs"] == 0 and e1b["rel_l2"] < 4e-3 and e2["rel_l2"] < 1e-5
    for (T, V) in [(300, 2048), (64, 49152), (17, 128256)]:
        logits = (torch.randn(T, V, device="cuda") * 2).bfloat16()
        labels = torch.randint(0, V, (T,), device="cuda")
        labels[::7] = -100
        lf = logits.float().requires_grad_(True)
        lr = torch.nn.functional.cross_entropy(lf, labels, ignore_index=-100)
        lr.backward()
        loss, loss_tok, dl = k.cross_entropy_fwd_bwd(logits.clone(), labels)
        torch.cuda.synchronize()
        el = abs(loss.item() - lr.item()) / abs(lr.item())
        eg = _err(dl, lf.grad)
        res[f"ce_{T}x{V}"] = {"loss_rel": el, "grad": eg}
        ok = ok and el < 1e-5 and eg["rel_l2"] < 5e-3
    T, V = 8192, 49152
    logits = torch.randn(T, V, device="cuda").bfloat16()
    labels = torch.randint(0, V, (T,), device="cuda")
    ms = _time(lambda: k.cross_entropy_fwd_bwd(logits, labels), iters=5)
    res["ce_GBps_alg_4TV"] = 4.0 * T * V / ms / 1e6
    res["ok"] = bool(ok)
    return res


@case
def optimizer():
    torch = _t()
    from dolomite_engine_b200 import kernels as k

    n = 1_000_003
    p = torch.randn(n, device="cuda")
    g = torch.randn(n, device="cuda") * 0.1
    p_ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([p_ref], lr=1e-3, betas=(0.9, 0.95), eps=1e-10, weight_decay=0.1)
    m = torch.zeros(n, device="cuda")
    v = torch.zeros(n, device="cuda")
    pb = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    ss = torch.zeros(1, device="cuda")
    coef = torch.empty(1, device="cuda")
    norm = torch.empty(1, device="cuda")
    k.sumsq_accum(g, ss)
    k.clip_coef(ss, 1.0, coef, norm)
    ref_norm = g.norm().item()
    for step in (1, 2, 3):
        p_ref.grad = g * min(1.0, 1.0 / (ref_norm + 1e-6))
        opt.step()
        k.adamw_step(p, g, m, v, pb, 1e-3, 0.9, 0.95, 1e-10, 0.1, step, clip=coef)
    e = _err(p, p_ref.data)
    eb = _err(pb, p_ref.data.bfloat16())
    # colsum + add_scaled
    x = torch.randn(777, 1024, device="cuda").bfloat16()
    out = torch.zeros(1024, device="cuda")
    k.colsum_accum(x, out)
    ec = _err(out, x.float().sum(0))
    a = torch.randn(4096, device="cuda").bfloat16()
    b = torch.randn(4096, device="cuda").bfloat16()
    ea = _err(k.add_scaled(a, b, 0.22), a + (b * 0.22))
    ok = e["rel_l2"] < 1e-5 and abs(norm.item() - ref_norm) / ref_norm < 1e-4 and ec["rel_l2"] < 1e-4 and ea["max_abs"] < 1e-9 + 0.04
    return {"adamw": e, "adamw_bf16": eb, "norm": [norm.item(), ref_norm], "colsum": ec, "add_scaled": ea, "ok": bool(ok)}


@case
def elementwise_bench_c2():
    """HBM-bound kernels at the C2 micro-batch-4 sizes (T=16384, H=2560, F=10240, 32 heads x hd 80): GB/s of algorithmic bytes"""
    torch = _t()
    from dolomite_engine_b200 import kernels as k

    T, H, F, nh, hd = 16384, 2560, 10240, 32, 80
    res = {}
    x = torch.randn(T, H, device="cuda").bfloat16()
    dy = torch.randn(T, H, device="cuda").bfloat16()
    dres = torch.randn(T, H, device="cuda").bfloat16()
    w = torch.ones(H, device="cuda").bfloat16()
    y = torch.empty_like(x)
    dx = torch.empty_like(x)
    dw = torch.zeros(H, device="cuda")
    _, rstd = k.rmsnorm_fwd(x, w, 1e-5, out=y)

    def rec(name, ms, nbytes):
        res[name] = {"ms": ms, "GBps": nbytes / ms / 1e6}

    rec("rmsnorm_fwd", _time(lambda: k.rmsnorm_fwd(x, w, 1e-5, out=y)), 4.0 * T * H)
    rec("rmsnorm_bwd_fused_residual", _time(lambda: k.rmsnorm_bwd(dy, x, w, rstd, dw, dx_add=dres, out=dx)), 8.0 * T * H)
    fc = torch.randn(T, 2 * F, device="cuda").bfloat16()
    act = torch.empty(T, F, device="cuda", dtype=torch.bfloat16)
    dact = torch.randn(T, F, device="cuda").bfloat16()
    dfc = torch.empty_like(fc)
    rec("swiglu_fwd", _time(lambda: k.swiglu_fwd(fc, out=act)), 6.0 * T * F)
    rec("swiglu_bwd", _time(lambda: k.swiglu_bwd(dact, fc, out=dfc)), 10.0 * T * F)
    qkv = torch.randn(T, 3 * H, device="cuda").bfloat16()
    npos = 4096
    cos = torch.randn(npos, hd, device="cuda").bfloat16()
    sin = torch.randn(npos, hd, device="cuda").bfloat16()
    pos = (torch.arange(T, device="cuda") % npos)
    rec("rope_qk", _time(lambda: k.rope_qk_inplace(qkv, nh, 1, hd, cos, sin, pos)), 4.0 * T * 2 * nh * hd)
    bias_g = torch.zeros(2 * F, device="cuda")
    rec("colsum_fc", _time(lambda: k.colsum_accum(dfc, bias_g)), 2.0 * T * 2 * F)
    n = 104_900_000
    p = torch.randn(n, device="cuda")
    g = torch.randn(n, device="cuda") * 0.01
    m = torch.zeros(n, device="cuda")
    v = torch.zeros(n, device="cuda")
    pb = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    ss = torch.zeros(1, device="cuda")
    rec("adamw_block", _time(lambda: k.adamw_step(p, g, m, v, pb, 1e-4, 0.9, 0.95, 1e-8, 0.1, 1), iters=10), 30.0 * n)
    rec("sumsq_block", _time(lambda: k.sumsq_accum(g, ss), iters=10), 4.0 * n)
    rec("zero_block(torch fill)", _time(lambda: g.zero_(), iters=10), 4.0 * n)
    res["ok"] = True
    return res


@case
def ce_bench():
    """row-resident cross entropy at the C2 / C5 / C4 vocabulary widths: GB/s of algorithmic bytes (4 B per logit)"""
    torch = _t()
    from dolomite_engine_b200 import kernels as k

    res = {}
    for name, (T, V) in {"c2_V49152": (8192, 49152), "c5_V128256": (4096, 128256), "c4_V50304": (8192, 50304)}.items():
        logits = torch.randn(T, V, device="cuda").bfloat16()
        labels = torch.randint(0, V, (T,), device="cuda")
        scratch = k.cross_entropy_count(labels)
        loss_tok = torch.empty(T, device="cuda")
        ms = _time(lambda: k.cross_entropy_rows(logits, labels, loss_tok, scratch), iters=10)
        res[name] = {"ms": ms, "GBps": 4.0 * T * V / ms / 1e6}
    res["ok"] = True
    return res


@case
def wgrad_multi_bench():
    """the four weight gradients of one C2 block at T = 24576 (bench micro-batch 6): four launches with the per-thread fp32
    epilogue (round 1), four launches with the TMA tile epilogue, ONE multi-problem launch, and cuBLAS (torch.matmul)"""
    torch = _t()
    from dolomite_engine_b200 import kernels as k

    T, H, F = 24576, 2560, 10240
    shapes = {"c_proj_mlp": (H, F), "c_fc": (2 * F, H), "c_proj_attn": (H, H), "c_attn": (3 * H, H)}
    g = torch.Generator(device="cuda").manual_seed(0)
    probs = []
    for M, N in shapes.values():
        dy = (torch.randn(T, M, device="cuda", generator=g) * 0.1).bfloat16()
        x = (torch.randn(T, N, device="cuda", generator=g) * 0.1).bfloat16()
        probs.append((dy, x, torch.zeros(M, N, device="cuda"), 1.0, True))
    flops = sum(2.0 * T * M * N for M, N in shapes.values())

    def separate(flags):
        def f():
            for dy, x, dw, a, _ in probs:
                k.gemm(dy, x, a_mn=True, b_mn=True, out=dw, c=dw, beta=1.0, flags=flags)
        return f

    def store(flags):
        def f():
            for dy, x, dw, a, _ in probs:
                k.gemm(dy, x, a_mn=True, b_mn=True, out=dw, flags=flags)
        return f

    def cublas():
        for dy, x, dw, _, _ in probs:
            torch.matmul(dy.t(), x)

    res = {}
    def multi(acc):
        return lambda: k.gemm_wgrad_multi([(a, b, c, d, acc) for a, b, c, d, _ in probs])

    for name, fn, tma in (("separate_direct_accumulate", separate(k.GEMM_DIRECT_EPILOGUE), 0),
                          ("separate_direct_overwrite", store(k.GEMM_DIRECT_EPILOGUE), 0),
                          ("separate_tma_reduce_add", separate(k.GEMM_F32_TMA_EPILOGUE), 0),
                          ("separate_tma_store_overwrite", store(k.GEMM_F32_TMA_EPILOGUE), 0),
                          ("multi_direct_accumulate", multi(True), 0), ("multi_direct_overwrite", multi(False), 0),
                          ("multi_tma_accumulate", multi(True), 1), ("multi_tma_overwrite", multi(False), 1),
                          ("cublas_bf16_out", cublas, 0), ("separate_direct_accumulate_again", separate(k.GEMM_DIRECT_EPILOGUE), 0)):
        k.set_option("gemm_f32_tma_epilogue", tma)
        ms = _time(fn, iters=10)
        res[name] = {"ms": ms, "tflops": flops / ms / 1e9}
    k.set_option("gemm_f32_tma_epilogue", 0)
    # correctness of the multi launch against torch on one problem
    dy, x, dw, _, _ = probs[2]
    dw.zero_()
    k.gemm_wgrad_multi([(p[0], p[1], p[2], 1.0, False) for p in probs])
    ref = dy.float().t() @ x.float()
    res["rel_l2_vs_torch"] = ((dw - ref).norm() / ref.norm()).item()
    res["ok"] = bool(res["rel_l2_vs_torch"] < 1e-2)
    return res


# ---------------------------------------------------------------------------------------------
def _attn_ref(qkv, cu, ng, g, hd, scale, dout=None):
    """fp32 reference: per-document causal softmax attention on the packed slot layout."""
    torch = _t()
    T = qkv.shape[0]
    x = qkv.float().view(T, ng, g + 2, hd).detach().clone().requires_grad_(True)
    q = x[:, :, :g].reshape(T, ng * g, hd)
    k = x[:, :, g].repeat_interleave(g, dim=1)
    v = x[:, :, g + 1].repeat_interleave(g, dim=1)
    outs, lses = [], []
    cu_l = cu.tolist()
    for d in range(len(cu_l) - 1):
        s, e = cu_l[d], cu_l[d + 1]
        if e == s:
            continue
        qd, kd, vd = q[s:e].transpose(0, 1), k[s:e].transpose(0, 1), v[s:e].transpose(0, 1)
        sc = torch.matmul(qd, kd.transpose(1, 2)) * scale
        mask = torch.ones(e - s, e - s, dtype=torch.bool, device=qkv.device).tril()
        sc = sc.masked_fill(~mask, float("-inf"))
        lses.append(torch.logsumexp(sc, dim=-1))
        outs.append(torch.matmul(torch.softmax(sc, dim=-1), vd).transpose(0, 1))
    out = torch.cat(outs, 0)
    lse = torch.cat(lses, 1)
    dx = None
    if dout is not None:
        out.backward(dout.float().view(T, ng * g, hd))
        dx = x.grad.view(T, -1)
    return out.reshape(T, -1).detach(), lse.detach(), dx


def _attn_case(lens, ng, g, hd, bwd=True, seed=5):
    torch = _t()
    from dolomite_engine_b200 import kernels as k

    gen = torch.Generator(device="cuda").manual_seed(seed)
    T = sum(lens)
    width = ng * (g + 2) * hd
    qkv = (torch.randn(T, width, device="cuda", generator=gen)).bfloat16()
    cu = torch.tensor([0] + list(__import__("itertools").accumulate(lens)), dtype=torch.int32, device="cuda")
    scale = hd ** -0.5
    dout = (torch.randn(T, ng * g * hd, device="cuda", generator=gen)).bfloat16()
    out, lse = k.attn_varlen_fwd(qkv, cu, max(lens), ng, g, hd, scale)
    torch.cuda.synchronize()
    ro, rl, rdx = _attn_ref(qkv, cu, ng, g, hd, scale, dout if bwd else None)
    res = {"fwd": _err(out, ro), "lse": _err(lse, rl), "T": T}
    ok = res["fwd"]["rel_l2"] < 1e-2 and res["lse"]["max_abs"] < 2e-2
    if bwd:
        dqkv = k.attn_varlen_bwd(dout, qkv, out, lse, cu, max(lens), ng, g, hd, scale)
        torch.cuda.synchronize()
        d = dqkv.float().view(T, ng, g + 2, hd)
        r = rdx.view(T, ng, g + 2, hd)
        res["dq"] = _err(d[:, :, :g], r[:, :, :g])
        res["dk"] = _err(d[:, :, g], r[:, :, g])
        res["dv"] = _err(d[:, :, g + 1], r[:, :, g + 1])
        ok = ok and all(res[x]["rel_l2"] < 2e-2 for x in ("dq", "dk", "dv"))
    res["ok"] = bool(ok)
    return res


@case
def attn_hd64_single_tile():
    return _attn_case([128], 2, 1, 64)


@case
def attn_hd64_two_tiles():
    return _attn_case([256], 2, 1, 64)


@case
def attn_hd64_ragged_docs():
    return _attn_case([100, 37, 300, 1, 129], 4, 1, 64)


@case
def attn_hd80_ragged_docs():
    return _attn_case([200, 130, 515], 4, 1, 80)


@case
def attn_hd128_gqa():
    return _attn_case([300, 77, 260], 2, 4, 128)


@case
def attn_hd32_mqa():
    return _attn_case([150, 250], 1, 4, 32)


@case
def attn_hd96():
    return _attn_case([384, 100], 2, 1, 96)


@case
def attn_hd16():
    return _attn_case([140], 2, 2, 16)


def _attn_bench(S, B, nh, hd):
    torch = _t()
    from dolomite_engine_b200 import kernels as k
    from flash_attn.flash_attn_interface import flash_attn_varlen_func

    T = S * B
    qkv = torch.randn(T, nh * 3 * hd, device="cuda").bfloat16()
    cu = torch.arange(0, T + 1, S, dtype=torch.int32, device="cuda")
    scale = hd ** -0.5
    dout = torch.randn(T, nh * hd, device="cuda").bfloat16()
    out, lse = k.attn_varlen_fwd(qkv, cu, S, nh, 1, hd, scale)
    v = qkv.view(T, nh, 3, hd)
    q, kk, vv = v[:, :, 0], v[:, :, 1], v[:, :, 2]
    fo = flash_attn_varlen_func(q, kk, vv, cu, cu, S, S, 0.0, softmax_scale=scale, causal=True)
    res = {"vs_flash_fwd": _err(out, fo.reshape(T, -1))}
    flops_fwd = 4.0 * S * S * hd * nh * B / 2
    ms = _time(lambda: k.attn_varlen_fwd(qkv, cu, S, nh, 1, hd, scale, out=out), iters=10)
    res["fwd_ms"] = ms
    res["fwd_tflops_causal"] = flops_fwd / ms / 1e9
    if hd >= 64:  # A/B: the single-buffer forward (two CTAs per SM) against the split-softmax forward
        k.set_option("attn_fwd_split", 2)
        out, lse = k.attn_varlen_fwd(qkv, cu, S, nh, 1, hd, scale)
        res["split_vs_flash_fwd"] = _err(out, fo.reshape(T, -1))
        ms2 = _time(lambda: k.attn_varlen_fwd(qkv, cu, S, nh, 1, hd, scale, out=out), iters=10)
        res["split_fwd_ms"] = ms2
        res["split_fwd_tflops_causal"] = flops_fwd / ms2 / 1e9
        k.set_option("attn_fwd_split", 0)
        out1, lse1 = k.attn_varlen_fwd(qkv, cu, S, nh, 1, hd, scale)
        res["single_buffer_vs_flash_fwd"] = _err(out1, fo.reshape(T, -1))
        res["lse_split_vs_single_max_abs"] = float((lse - lse1).abs().max())
        ms1 = _time(lambda: k.attn_varlen_fwd(qkv, cu, S, nh, 1, hd, scale, out=out1), iters=10)
        res["single_buffer_fwd_ms"] = ms1
        res["single_buffer_fwd_tflops_causal"] = flops_fwd / ms1 / 1e9
        k.set_option("attn_fwd_split", 1)
    ms_f = _time(lambda: flash_attn_varlen_func(q, kk, vv, cu, cu, S, S, 0.0, softmax_scale=scale, causal=True), iters=10)
    res["flash_fwd_ms"] = ms_f
    res["flash_fwd_tflops_causal"] = flops_fwd / ms_f / 1e9
    dqkv = torch.empty_like(qkv)
    k.attn_varlen_bwd(dout, qkv, out, lse, cu, S, nh, 1, hd, scale, dqkv=dqkv)
    qf = q.detach().clone().requires_grad_(True)
    kf = kk.detach().clone().requires_grad_(True)
    vf = vv.detach().clone().requires_grad_(True)
    fo2 = flash_attn_varlen_func(qf, kf, vf, cu, cu, S, S, 0.0, softmax_scale=scale, causal=True)
    fo2.backward(dout.view(T, nh, hd))
    d = dqkv.view(T, nh, 3, hd)
    res["vs_flash_dq"] = _err(d[:, :, 0], qf.grad)
    res["vs_flash_dk"] = _err(d[:, :, 1], kf.grad)
    res["vs_flash_dv"] = _err(d[:, :, 2], vf.grad)
    ms_b = _time(lambda: k.attn_varlen_bwd(dout, qkv, out, lse, cu, S, nh, 1, hd, scale, dqkv=dqkv), iters=5)
    res["bwd_ms"] = ms_b
    res["bwd_tflops_causal"] = 2.5 * flops_fwd / ms_b / 1e9

    def fb():
        o = flash_attn_varlen_func(qf, kf, vf, cu, cu, S, S, 0.0, softmax_scale=scale, causal=True)
        o.backward(dout.view(T, nh, hd))

    ms_fb = _time(fb, iters=5)
    res["flash_fwd_bwd_ms"] = ms_fb
    res["ok"] = bool(res["vs_flash_fwd"]["rel_l2"] < 1e-2 and res["vs_flash_dq"]["rel_l2"] < 2e-2 and res["vs_flash_dk"]["rel_l2"] < 2e-2 and res["vs_flash_dv"]["rel_l2"] < 2e-2)
    return res


def _attn_order_ab(S, B, ng, g, hd, rounds=3, orders=(0, 8, 1024)):
    """CTA order of the attention kernels (`attn_head_fastest`): round 1's tiles fastest (0), chunks of 8 heads (default) and
    all heads in one chunk (1024); forward and backward, interleaved; the forward must be bit-identical (the order changes
    nothing but timing -- the order of the backward's dQ reduction aside)"""
    torch = _t()
    from dolomite_engine_b200 import kernels as k

    T, nh = S * B, ng * g
    qkv = torch.randn(T, ng * (g + 2) * hd, device="cuda").bfloat16()
    cu = torch.arange(0, T + 1, S, dtype=torch.int32, device="cuda")
    scale = hd ** -0.5
    dout = torch.randn(T, nh * hd, device="cuda").bfloat16()
    flops_fwd = 4.0 * S * S * hd * nh * B / 2
    res = {"shape": [S, B, ng, g, hd]}
    default = k.get_option("attn_head_fastest")
    outs = {}
    tf = {o: [] for o in orders}
    tb = {o: [] for o in orders}
    try:
        for _ in range(rounds):
            for order in orders:
                k.set_option("attn_head_fastest", order)
                out, lse = k.attn_varlen_fwd(qkv, cu, S, ng, g, hd, scale)
                dqkv = k.attn_varlen_bwd(dout, qkv, out, lse, cu, S, ng, g, hd, scale)
                outs[order] = (out.clone(), dqkv.clone())
                tf[order].append(_time(lambda: k.attn_varlen_fwd(qkv, cu, S, ng, g, hd, scale, out=out), iters=10))
                tb[order].append(_time(lambda: k.attn_varlen_bwd(dout, qkv, out, lse, cu, S, ng, g, hd, scale, dqkv=dqkv), iters=5))
    finally:
        k.set_option("attn_head_fastest", default)
    res["fwd_bit_identical"] = bool(all(torch.equal(outs[orders[0]][0], outs[o][0]) for o in orders))
    res["bwd_rel_l2_between_orders"] = max(_err(outs[o][1], outs[orders[0]][1])["rel_l2"] for o in orders)
    for order in orders:
        res[f"order{order}"] = {"fwd_ms": round(min(tf[order]), 4), "bwd_ms": round(min(tb[order]), 4),
                                "fwd_tflops_causal": round(flops_fwd / min(tf[order]) / 1e9),
                                "bwd_tflops_causal": round(2.5 * flops_fwd / min(tb[order]) / 1e9)}
    res["ok"] = bool(res["fwd_bit_identical"] and res["bwd_rel_l2_between_orders"] < 1e-3)
    return res


def _attn_fwd_kernels_ab(S, B, ng, g, hd, rounds=3, splits=(0, 2, 3)):
    """forward kernels of one head_dim, interleaved: single buffer (attn_fwd_split = 0; two CTAs per SM at head_dim <= 80),
    split softmax with two (2) and with four (3) threads per query row"""
    torch = _t()
    from dolomite_engine_b200 import kernels as k

    T, nh = S * B, ng * g
    qkv = torch.randn(T, ng * (g + 2) * hd, device="cuda").bfloat16()
    cu = torch.arange(0, T + 1, S, dtype=torch.int32, device="cuda")
    scale = hd ** -0.5
    flops_fwd = 4.0 * S * S * hd * nh * B / 2
    res = {"shape": [S, B, ng, g, hd]}
    default = k.get_option("attn_fwd_split")
    outs, ts = {}, {sp: [] for sp in splits}
    try:
        for _ in range(rounds):
            for sp in splits:
                k.set_option("attn_fwd_split", sp)
                out, lse = k.attn_varlen_fwd(qkv, cu, S, ng, g, hd, scale)
                outs[sp] = (out.clone(), lse.clone())
                ts[sp].append(_time(lambda: k.attn_varlen_fwd(qkv, cu, S, ng, g, hd, scale, out=out), iters=10))
    finally:
        k.set_option("attn_fwd_split", default)
    for sp in splits:
        res[f"split{sp}"] = {"fwd_ms": round(min(ts[sp]), 4), "fwd_tflops_causal": round(flops_fwd / min(ts[sp]) / 1e9),
                             "rel_l2_vs_split0": _err(outs[sp][0], outs[splits[0]][0])["rel_l2"],
                             "lse_max_abs_vs_split0": float((outs[sp][1] - outs[splits[0]][1]).abs().max())}
    res["ok"] = bool(all(res[f"split{sp}"]["rel_l2_vs_split0"] < 5e-3 for sp in splits))
    return res


@case
def attn_fwd_kernels_c2():
    return _attn_fwd_kernels_ab(4096, 6, 32, 1, 80)


@case
def attn_fwd_kernels_c5():
    return _attn_fwd_kernels_ab(8192, 1, 8, 4, 128)


@case
def attn_fwd_kernels_c4():
    return _attn_fwd_kernels_ab(2048, 8, 16, 1, 128)


@case
def attn_order_c2():
    return _attn_order_ab(4096, 6, 32, 1, 80)


@case
def attn_order_c5():
    return _attn_order_ab(8192, 1, 8, 4, 128)


@case
def attn_order_c4():
    return _attn_order_ab(2048, 8, 16, 1, 128)


def _attn_bwd_variants(S, B, nh, hd, rounds=3):
    """A/B/C of the pipelined backward's softmax-warp variants, interleaved so that clock / power drift hits all alike"""
    torch = _t()
    from dolomite_engine_b200 import kernels as k
    from flash_attn.flash_attn_interface import flash_attn_varlen_func

    T = S * B
    qkv = torch.randn(T, nh * 3 * hd, device="cuda").bfloat16()
    cu = torch.arange(0, T + 1, S, dtype=torch.int32, device="cuda")
    scale = hd ** -0.5
    dout = torch.randn(T, nh * hd, device="cuda").bfloat16()
    out, lse = k.attn_varlen_fwd(qkv, cu, S, nh, 1, hd, scale)
    v = qkv.view(T, nh, 3, hd)
    qf, kf, vf = (v[:, :, i].detach().clone().requires_grad_(True) for i in range(3))
    fo = flash_attn_varlen_func(qf, kf, vf, cu, cu, S, S, 0.0, softmax_scale=scale, causal=True)
    fo.backward(dout.view(T, nh, hd))
    flops = 2.5 * 4.0 * S * S * hd * nh * B / 2
    res = {"ok": True}
    dqkv = torch.empty_like(qkv)
    times = {0: [], 1: [], 2: []}
    try:
        for variant in (0, 1, 2):
            k.set_option("attn_bwd_variant", variant)
            k.attn_varlen_bwd(dout, qkv, out, lse, cu, S, nh, 1, hd, scale, dqkv=dqkv)
            d = dqkv.view(T, nh, 3, hd)
            e = {"dq": _err(d[:, :, 0], qf.grad), "dk": _err(d[:, :, 1], kf.grad), "dv": _err(d[:, :, 2], vf.grad)}
            res[f"variant{variant}_err"] = {n: x["rel_l2"] for n, x in e.items()}
            res["ok"] = res["ok"] and all(x["rel_l2"] < 2e-2 for x in e.values())
        for _ in range(rounds):
            for variant in (0, 1, 2):
                k.set_option("attn_bwd_variant", variant)
                times[variant].append(_time(lambda: k.attn_varlen_bwd(dout, qkv, out, lse, cu, S, nh, 1, hd, scale, dqkv=dqkv), iters=5))
    finally:
        k.set_option("attn_bwd_variant", 2)
    for variant, ts in times.items():
        res[f"variant{variant}_ms"] = ts
        res[f"variant{variant}_tflops_causal"] = flops / min(ts) / 1e9
    return res


def _attn_bwd_ablation(S, B, nh, hd):
    """TIMING ONLY (results are wrong by construction): which stage bounds the pipelined backward?  Bits of `attn_bwd_ablate`:
    1 no TMA reduce-add of dQ, 2 no dQ drain at all, 4 no dQ MMAs, 8 no softmax math, 16 no dV / dK MMAs, 32 no dS^T smem stores"""
    torch = _t()
    from dolomite_engine_b200 import kernels as k

    T = S * B
    qkv = torch.randn(T, nh * 3 * hd, device="cuda").bfloat16()
    cu = torch.arange(0, T + 1, S, dtype=torch.int32, device="cuda")
    scale = hd ** -0.5
    dout = torch.randn(T, nh * hd, device="cuda").bfloat16()
    out, lse = k.attn_varlen_fwd(qkv, cu, S, nh, 1, hd, scale)
    dqkv = torch.empty_like(qkv)
    flops = 2.5 * 4.0 * S * S * hd * nh * B / 2
    res = {"ok": True}
    masks = [0, 1, 3, 7, 8, 32, 40, 16, 47, 63]
    times = {m: [] for m in masks}
    try:
        for _ in range(2):
            for m in masks:
                k.set_option("attn_bwd_ablate", m)
                times[m].append(_time(lambda: k.attn_varlen_bwd(dout, qkv, out, lse, cu, S, nh, 1, hd, scale, dqkv=dqkv), iters=5))
    finally:
        k.set_option("attn_bwd_ablate", 0)
    for m, ts in times.items():
        res[f"ablate{m}_ms"] = min(ts)
    # the small kernels around the main one (delta, memset, finalize) are inside these times: measure them alone
    return res


@case
def attn_bwd_trace_hd80():
    """clock64 timeline of ONE CTA of the pipelined backward (debug hook, not part of the ABI): gpurun_out/bwd_trace_*.npy"""
    import ctypes

    import numpy as np

    torch = _t()
    from dolomite_engine_b200 import _lib
    from dolomite_engine_b200 import kernels as k

    S, B, nh, hd = 4096, 2, 32, 80
    T = S * B
    qkv = torch.randn(T, nh * 3 * hd, device="cuda").bfloat16()
    cu = torch.arange(0, T + 1, S, dtype=torch.int32, device="cuda")
    scale = hd ** -0.5
    dout = torch.randn(T, nh * hd, device="cuda").bfloat16()
    out, lse = k.attn_varlen_fwd(qkv, cu, S, nh, 1, hd, scale)
    dqkv = torch.empty_like(qkv)
    lib = _lib.load()
    fn = lib.dolomite_b200_debug_attn_bwd_trace
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
    fn.restype = ctypes.c_int
    res = {"ok": True}
    os.makedirs("gpurun_out", exist_ok=True)
    try:
        for ablate in (0, 7, 63):
            k.set_option("attn_bwd_ablate", ablate)
            for _ in range(3):
                k.attn_varlen_bwd(dout, qkv, out, lse, cu, S, nh, 1, hd, scale, dqkv=dqkv)
            buf = torch.zeros(5 * 2048, dtype=torch.int64, device="cuda")
            fn(buf.data_ptr(), 8)
            k.attn_varlen_bwd(dout, qkv, out, lse, cu, S, nh, 1, hd, scale, dqkv=dqkv)
            torch.cuda.synchronize()
            fn(None, 0)
            np.save(f"gpurun_out/bwd_trace_ablate{ablate}.npy", buf.cpu().numpy())
            res[f"events_ablate{ablate}"] = int((buf != 0).sum())
    finally:
        fn(None, 0)
        k.set_option("attn_bwd_ablate", 0)
    return res


@case
def attn_bwd_ablation_hd80():
    return _attn_bwd_ablation(4096, 2, 32, 80)


@case
def attn_bwd_ablation_hd64():
    return _attn_bwd_ablation(4096, 2, 32, 64)


@case
def attn_bwd_variants_hd80():
    return _attn_bwd_variants(4096, 2, 32, 80)


@case
def attn_bwd_variants_hd64():
    return _attn_bwd_variants(4096, 2, 32, 64)


@case
def attn_bench_c2():
    return _attn_bench(4096, 2, 32, 80)


@case
def attn_bench_hd128_s4096():
    return _attn_bench(4096, 2, 16, 128)


@case
def env_info():
    torch = _t()
    import subprocess as sp

    smi = sp.run(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total", "--format=csv"], capture_output=True, text=True).stdout
    return {"torch": torch.__version__, "device": torch.cuda.get_device_name(0), "cpu_count": os.cpu_count(), "smi": smi, "ok": True}


# ---------------------------------------------------------------------------------------------
@case
def moe_bench_c4():
    """MoEDolomite C4 shape (H 2048, 16 heads hd 128, 8 experts top-2, F 4096 per expert, seq 2048), 2 layers, mbs 8
    (T = 16384): fwd+bwd step time and the rate of the grouped expert GEMMs (CUDA events around every launch)."""
    torch = _t()
    import numpy as np

    from dolomite_engine_b200 import kernels as k
    from dolomite_engine_b200.hf_models import MoEDolomiteConfig, MoEDolomiteForCausalLM

    H, E, topk, F, S, mbs, L, V = 2048, 8, 2, 4096, 2048, 8, 2, 50304
    cfg = MoEDolomiteConfig(vocab_size=V, n_positions=S, n_embd=H, n_layer=L, n_head=16, n_inner=F,
                            attention_head_type="mha", add_bias=False, num_experts=E, num_experts_per_tok=topk,
                            position_embedding_type="rope", normalization_function="rmsnorm",
                            activation_function="swiglu", resid_pdrop=0, embd_pdrop=0, attn_pdrop=0, eos_token_id=0)
    model = MoEDolomiteForCausalLM(cfg, seed=1)
    model.assume_unit_loss_grad = True
    T = S * mbs
    g = torch.Generator(device="cuda").manual_seed(0)
    ids = torch.randint(0, V, (T,), device="cuda", generator=g)
    labels = torch.randint(0, V, (T,), device="cuda", generator=g)
    pos = (torch.arange(T, device="cuda") % S)
    cu = torch.arange(0, T + 1, S, dtype=torch.int32, device="cuda")

    def step():
        model.engine.zero_grad()
        loss = model.forward_pretraining_loss(ids, pos, cu, S, labels)
        loss.backward()
        return loss

    # time every grouped GEMM launch of one step
    import dolomite_engine_b200._lib as L_
    records = []
    orig_call = L_.call

    def timed_call(name, *a):
        if "grouped" in name:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = orig_call(name, *a)
            e1.record()
            records.append((name, e0, e1))
            return r
        return orig_call(name, *a)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    L_.call = timed_call
    k._lib.call = timed_call
    try:
        loss = step()
        torch.cuda.synchronize()
    finally:
        L_.call = orig_call
        k._lib.call = orig_call
    grouped_ms = sum(e0.elapsed_time(e1) for _, e0, e1 in records)
    # expert FLOPs of one step: fwd 2 GEMMs + bwd 4 (dgrad + wgrad each); T*topk routed rows; c_fc 2F (glu), c_proj F
    flops = L * 3 * 2.0 * (T * topk) * H * (2 * F + F)
    ms = _time(step, iters=5)
    return {"loss": float(loss.item()), "step_ms_fwd_bwd": ms, "tokens_per_s_fwd_bwd": T / ms * 1e3,
            "grouped_gemm_launches": len(records), "grouped_gemm_ms": grouped_ms,
            "grouped_gemm_tflops": flops / grouped_ms / 1e9, "ok": bool(np.isfinite(loss.item()))}


@case
def moe_layer_fwd_c4():
    """only the MoE MLP forward at the C4 shape (for ncu: launch order per iteration = router gemm, route kernels, gather,
    grouped c_fc gemm, swiglu, grouped c_proj gemm, combine)"""
    torch = _t()
    from dolomite_engine_b200 import kernels as k

    H, E, topk, F, T = 2048, 8, 2, 4096, 16384
    x = (torch.randn(T, H, device="cuda") * 0.5).bfloat16()
    gate = (torch.randn(E, H, device="cuda") * 0.05).bfloat16()
    w_fc = (torch.randn(E, 2 * F, H, device="cuda") * 0.02).bfloat16()
    w_proj = (torch.randn(E, H, F, device="cuda") * 0.02).bfloat16()
    res = torch.zeros(T, H, device="cuda", dtype=torch.bfloat16)

    def make(fused: bool, flags):
        def fwd():
            logits = k.gemm(x, gate, flags=0)
            plan = k.moe_route(logits, topk)
            if fused:
                fc = k.gemm_grouped_m_gather(x, w_fc, plan, flags=flags)
            else:
                fc = k.gemm_grouped_m(k.moe_gather(x, plan), w_fc, plan, b_mn=False, flags=flags)
            act = k.swiglu_fwd(fc)
            yg = k.gemm_grouped_m(act, w_proj, plan, b_mn=False, flags=flags)
            return k.moe_combine(yg, plan, c=res, alpha=1.0)
        return fwd

    flops = 2.0 * T * topk * H * 3 * F
    out = {}
    for name, fused, flags in (("fused_gather_pair", True, k.GEMM_TMA_STORE), ("gather_kernel_pair", False, k.GEMM_TMA_STORE),
                               ("gather_kernel_single_cta", False, k.GEMM_TMA_STORE | 8)):
        ms_i = _time(make(fused, flags), iters=5, warmup=2)
        out[name] = {"fwd_ms": ms_i, "expert_gemm_tflops_incl_routing_kernels": flops / ms_i / 1e9}
    ms = out["fused_gather_pair"]["fwd_ms"]
    out["ok"] = True
    return out


@case
def attn_bench_hd128():
    """Llama-3-8B-like attention core (C5): S=8192, 32 heads, hd 128 (MHA layout of the packed slots)"""
    return _attn_bench(8192, 1, 32, 128)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case")
    ap.add_argument("--only", default="")
    ap.add_argument("--timeout", type=int, default=120)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "probe.jsonl"))
    a = ap.parse_args()
    if a.case:
        t0 = time.time()
        try:
            res = CASES[a.case]()
        except Exception as e:  # noqa
            import traceback

            res = {"ok": False, "error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc()[-1500:]}
        res["case"] = a.case
        res["wall_s"] = round(time.time() - t0, 2)
        print("PROBE_RESULT " + json.dumps(res))
        return
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    names = [n for n in CASES if a.only in n]
    with open(a.out, "a") as f:
        for n in names:
            try:
                proc = subprocess.run([sys.executable, __file__, "--case", n], capture_output=True, text=True, timeout=a.timeout)
                line = [l for l in proc.stdout.splitlines() if l.startswith("PROBE_RESULT ")]
                if line:
                    res = json.loads(line[-1][len("PROBE_RESULT "):])
                else:
                    res = {"case": n, "ok": False, "error": "no result", "rc": proc.returncode, "stdout": proc.stdout[-1500:], "stderr": proc.stderr[-1500:]}
            except subprocess.TimeoutExpired as e:
                res = {"case": n, "ok": False, "error": "timeout", "stdout": (e.stdout or b"")[-1000:].decode(errors="replace") if isinstance(e.stdout, bytes) else str(e.stdout)[-1000:]}
            f.write(json.dumps(res) + "\n")
            f.flush()
            brief = {k: v for k, v in res.items() if k in ("case", "ok", "error", "rel_l2", "max_abs", "tflops", "cublas_tflops", "ms")}
            print(json.dumps(brief), flush=True)


if __name__ == "__main__":
    main()
