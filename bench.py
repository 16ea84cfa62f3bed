"""Benchmark of the hot path: one full training step (fwd + bwd + reduce-scatter + clip + AdamW) of the
GPTDolomite Granite-3B-code shape (BASELINE.json configs[1]: 32L / 2560d / 32 heads hd=80 / F=10240 / V=49152, bf16,
seq 4096 padding-free, synthetic packed tokens) on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mbs 2] [--layers 32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line on rank 0 (contract in the task statement):
  value      whole-job tokens/s with the batch already resident in HBM (device-timed, max over ranks)
  e2e        the same metric through the reference-facing wrapper call `model({"text": cpu_tensor})` with pinned
             HOST buffers: H2D of the step's tokens and D2H of the loss inside the timed region
  roofline   dominant kernel = the tcgen05 GEMM: algorithmic FLOPs of every GEMM launch / CUDA-event time of that
             launch, measured live inside the timed region, against the measured bf16 peak (MEASURED_PEAKS.json)
  cpu_baseline  the oracle (CPU restatement of the reference, torch-eager fp32, eager attention) timed on the host
             cores on a bounded sample (rank 0, N=1 only)
`--impl reference` times that CPU implementation as its own arm (all host threads), same metric/config.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C2 = dict(model_type="gpt_dolomite", vocab_size=49152, n_positions=4096, n_embd=2560, n_layer=32, n_head=32, n_inner=10240,
          attention_head_type="mha", position_embedding_type="rope", activation_function="swiglu",
          normalization_function="rmsnorm", layer_norm_epsilon=1e-5, add_bias=True, resid_pdrop=0, embd_pdrop=0,
          attn_pdrop=0, upcast_logits_for_loss=True, eos_token_id=0)
SEQ = 4096


def flops_per_token(cfg: dict, seq: int) -> float:
    """reference FLOP model, train_utils.py:197-236 (full SxS attention, no causal discount)"""
    h, f, n, L, v = cfg["n_embd"], cfg["n_inner"], cfg["n_head"], cfg["n_layer"], cfg["vocab_size"]
    k = cfg.get("num_key_value_heads") or n
    return 3 * L * (4 * h * (h * (1 + k / n) + seq) + 6 * h * f) + 6 * h * v


def measured_peaks() -> dict:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["source"] = "measured"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows: list[list[str]] = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (reference restatement) on the host cores, bounded sample, extrapolated per layer
# ------------------------------------------------------------------------------------------------
def cpu_reference_tokens_per_s(cfg: dict, seq: int, steps: int = 1, warmup: int = 0, sample_layers: int = 1, sample_seq: int = 512):
    import numpy as np
    import torch

    import oracle.dolomite_oracle as O

    # "all the host threads it can use": torch's intra-op pool does not always scale to every core of a many-core
    # host, so pick the fastest thread count with a short GEMM probe and report it
    avail = os.cpu_count() or 1
    best, cores = None, avail
    a = torch.randn(2048, 2560)
    b = torch.randn(2560, 2560)
    for n in sorted({min(avail, c) for c in (16, 32, 64, avail)}):
        torch.set_num_threads(n)
        torch.mm(a, b)
        t0 = time.perf_counter()
        for _ in range(3):
            torch.mm(a, b)
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, cores = dt, n
    torch.set_num_threads(cores)
    small ={k: v for k, v in cfg.items() if k in O.OracleConfig.__dataclass_fields__}
    small.update(n_layer=sample_layers, n_positions=max(sample_seq, 16))
    ocfg = O.OracleConfig(**small)
    params = {k: v.requires_grad_(True) for k, v in O.init_params(ocfg, seed=1).items()}
    rng = np.random.default_rng(0)
    tokens = rng.integers(0, ocfg.vocab_size, size=(1, sample_seq + 1), dtype=np.int64)

    def one(n_layer_cfg):
        t0 = time.perf_counter()
        loss, _ = O.pretraining_loss(params, n_layer_cfg, tokens)
        loss.backward()
        return time.perf_counter() - t0

    for _ in range(max(warmup, 1)):
        one(ocfg)
    times = []
    for _ in range(max(steps, 1)):
        t_full = one(ocfg)
        # head/embedding-only cost: zero layers
        import dataclasses

        t_head = one(dataclasses.replace(ocfg, n_layer=0))
        times.append((t_full, t_head))
    t_full = sum(t[0] for t in times) / len(times)
    t_head = sum(t[1] for t in times) / len(times)
    t_layer = max(t_full - t_head, 1e-9) / sample_layers
    # attention cost grows with seq: per-token attention work at S is S/sample_seq times the sample's; the GEMM part
    # is per token.  Keep the estimate conservative (favourable to the CPU): scale only by layer count.
    t_token_full = (cfg["n_layer"] * t_layer + t_head) / sample_seq
    return 1.0 / t_token_full, {
        "value": 1.0 / t_token_full, "unit": "tokens/s", "cores": cores, "kind": "port",
        "sample": (f"oracle (torch-eager fp32, eager attention) fwd+bwd of {sample_layers} C2 block(s) + LM head at "
                   f"S={sample_seq}, mbs=1 ({t_full:.2f}s; head-only {t_head:.2f}s), extrapolated to {cfg['n_layer']} layers; "
                   "attention S^2 growth to S=4096 and the optimizer step are NOT charged (favours the CPU)"),
    }


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cfg = dict(C2)
    cfg["n_layer"] = args.layers
    t0 = time.perf_counter()
    tps, cb = cpu_reference_tokens_per_s(cfg, SEQ, steps=args.steps, warmup=min(args.warmup, 1))
    tokens_per_step = args.mbs * SEQ
    line = {
        "impl": "reference", "metric": "tokens_per_sec", "value": tps, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tokens_per_step / tps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "GPTDolomite Granite-3B-code shape (32L/2560d/hd80/F10240/V49152) bf16 seq4096 padding-free, "
                               "full train step", "micro_batch_size": args.mbs, "seq_len": SEQ, "n_layer": cfg["n_layer"]},
        "cpu_baseline": cb,
        "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from dolomite_engine_b200 import _lib
    from dolomite_engine_b200 import kernels as K
    from dolomite_engine_b200.distributed import ShardedDataParallel
    from dolomite_engine_b200.model_wrapper import ModelWrapperForPretraining
    from dolomite_engine_b200.optimization import get_optimizer
    from dolomite_engine_b200.pretrain import SyntheticPackedDataset
    from dolomite_engine_b200.train_utils import train_step

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        from dolomite_engine_b200.distributed import configure_comm_ctas

        configure_comm_ctas()
        dist.init_process_group(backend="nccl", rank=rank, world_size=world, device_id=dev)
    cfg = dict(C2)
    cfg["n_layer"] = args.layers
    mbs, seq = args.mbs, SEQ
    wrapper = ModelWrapperForPretraining(pretrained_config=cfg, micro_batch_size=mbs, sequence_length=seq, device=dev,
                                         world_size=world, rank=rank, init_on_device=True,
                                         reset_attention_mask=args.ragged, reset_position_ids=args.ragged)
    model = ShardedDataParallel(wrapper, dist.group.WORLD if world > 1 else None,
                                communication_dtype={"bf16": torch.bfloat16, "fp32": torch.float32}[args.comm_dtype])
    opt = get_optimizer("DolomiteFusedAdamW", {"lr": 1e-5, "weight_decay": 0.1, "betas": [0.9, 0.95], "eps": 1e-10}, model)
    data = SyntheticPackedDataset(cfg["vocab_size"], mbs, seq, rank=rank, eos_token_id=cfg["eos_token_id"], ragged=args.ragged)
    tokens_per_step = mbs * seq * world

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- leg 1: batch resident in HBM ----------------
    engine = wrapper.model.engine
    batch = next(data)["text"]
    ids, labels, pos, cu, max_seqlen = wrapper._stage(batch)
    ids, labels, pos, cu = ids.clone(), labels.clone(), pos.clone(), cu.clone()

    def resident_step():
        model.zero_grad()
        model._refresh_parameters_if_needed()
        loss = wrapper.model.forward_pretraining_loss(ids, pos, cu, max_seqlen, labels)
        loss.backward()
        model.clip_grad_norm_(1.0, fuse_into_optimizer=True)
        opt.step()
        return loss

    for _ in range(args.warmup):
        resident_step()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.reset_launch_counts()
    K.gemm_timer = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss_hist = []
    for _ in range(args.steps):
        loss = resident_step()
        loss_hist.append(loss.detach())
    e1.record()
    sync_all()
    ms_resident = e0.elapsed_time(e1) / args.steps
    launches = _lib.total_kernel_launches()
    gemm_records = K.gemm_timer
    K.gemm_timer = None
    gemm_flops = sum(r[0] for r in gemm_records)
    gemm_ms = sum(r[1].elapsed_time(r[2]) for r in gemm_records)
    last_loss = float(loss.item())

    if args.profile_step and rank == 0:
        # untimed diagnostic: per-kernel device time of ONE warm in-situ step (CUPTI via torch.profiler) -> JSON.
        # Unlike the ncu launch list (cold caches, serialised, base clocks) this is the step as it actually runs.
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            resident_step()
            torch.cuda.synchronize()
        agg, first, last = {}, None, None
        for ev in prof.events():
            if ev.device_type.name != "CUDA" or ev.device_time_total <= 0:
                continue
            a = agg.setdefault(ev.name[:120], [0, 0.0])
            a[0] += 1
            a[1] += ev.device_time_total / 1e3
            s0, s1 = ev.time_range.start, ev.time_range.end
            first = s0 if first is None else min(first, s0)
            last = s1 if last is None else max(last, s1)
        busy = sum(v[1] for v in agg.values())
        with open(args.profile_step, "w") as f:
            json.dump({"span_ms": (last - first) / 1e3, "sum_kernel_ms": busy, "ms_per_step_timed": ms_resident,
                       "kernels": sorted(([k, v[0], v[1]] for k, v in agg.items()), key=lambda r: -r[2])}, f, indent=1)

    # ---------------- leg 2: end to end through the wrapper call with host buffers ----------------
    def e2e_step():
        l, gn = train_step(model, opt, None, train_dataloader=data, gradient_accumulation_steps=1, gradient_clipping=1.0)
        return l

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    sync_all()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        e2e_step()  # includes H2D of the tokens (pinned) and D2H of loss / grad-norm (.item())
    t1.record()
    sync_all()
    ms_e2e = t0.elapsed_time(t1) / args.steps
    clocks = sampler.stop() if rank == 0 else None

    # max over ranks
    t = torch.tensor([ms_resident, ms_e2e], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_resident, ms_e2e = t.tolist()

    if rank == 0:
        peaks = measured_peaks()
        fpt = flops_per_token(cfg, seq)
        value = tokens_per_step / (ms_resident / 1e3)
        e2e_val = tokens_per_step / (ms_e2e / 1e3)
        gemm_tflops = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        peak = peaks["bf16_tflops_sustained"]
        line = {
            "metric": "tokens_per_sec", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_resident, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {
                "workload": "GPTDolomite Granite-3B-code shape (32L/2560d/hd80/F10240/V49152, 3.48B params) bf16 seq4096 "
                            "padding-free, full train step (fwd+bwd+reduce-scatter+clip+AdamW)",
                "micro_batch_size": mbs, "global_batch": mbs * world, "seq_len": seq, "n_layer": cfg["n_layer"],
                "parallelism": f"flat-bucket sharded data parallel x{world}", "packing": "ragged" if args.ragged else "uniform",
                "communication_dtype": args.comm_dtype,
                "l2": "working set (>=7 GB parameters + activations per step) far exceeds the 126 MB L2; no explicit flush",
            },
            "tokens_per_sec_per_gpu": value / world,
            "model_tflops_per_gpu": fpt * value / world / 1e12,
            "model_flops_per_token": fpt,
            "pct_of_bf16_peak_measured_sustained": 100.0 * fpt * value / world / 1e12 / peak,
            "pct_of_bf16_peak_measured_burst": 100.0 * fpt * value / world / 1e12 / peaks["bf16_tflops"],
            "loss": last_loss, "loss_history_same_batch": [float(x) for x in torch.stack(loss_hist).tolist()],
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "tokens/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": wrapper.h2d_bytes_per_step, "d2h_bytes_per_step": 8},
            "gpu_launches": launches,
            "peak_hbm_gb": torch.cuda.max_memory_allocated(dev) / 1e9,
            "roofline": {
                "bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05, all nn.Linear fwd/dgrad/wgrad + LM head)",
                "achieved": gemm_tflops, "peak": peak, "unit": "TFLOP/s",
                "frac": (gemm_tflops / peak) if gemm_tflops else None,
                # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from the committed `ncu --set full` capture
                # (c_fc forward, M=8192 N=20480 K=2560, CTA-pair kernel): 256.7 MB + 304.8 MB vs 482.3 MB algorithmic
                "traffic": 561486336, "traffic_algorithmic": 482344960,
                "traffic_source": "profiles/r01_ncu_gemm_fc_pair_call14.txt (tensor pipe 93.4 % active)",
                "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a long step)",
                "launches_timed": len(gemm_records), "share_of_step": gemm_ms / (ms_resident * args.steps),
            },
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                _, cb = cpu_reference_tokens_per_s(cfg, seq, steps=1, warmup=1)
                line["cpu_baseline"] = cb
            except Exception as e:  # the oracle is test infrastructure; never let it break the measurement
                line["cpu_baseline"] = {"value": None, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {type(e).__name__}: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mbs", type=int, default=6,
                    help="sequences of 4096 tokens per GPU per step, chosen to fill memory (SURVEY 8d): 4 -> 117 GB, 6 -> ~146 GB "
                         "of 180 GB; measured 41.1 k / 42.1 k / 42.3 k tokens/s at 4 / 6 / 7")
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--ragged", action="store_true")
    ap.add_argument("--comm-dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-step", default=None, metavar="JSON",
                    help="diagnostic: write per-kernel device times of one extra (untimed) step to this file")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
