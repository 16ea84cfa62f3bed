"""Benchmark of the hot path: one full training step (fwd + bwd + reduce-scatter + clip + AdamW) on N GPUs of one node.

    python bench.py [--config c2|c4|c5] [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|gpu_reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads (BASELINE.json `configs`; shapes pinned by SURVEY.md section 8d):
  c2 (default, the configuration the metric is quoted on)  GPTDolomite Granite-3B-code shape, 32L / 2560d / 32 heads hd 80 /
      F 10240 / V 49152, bf16, seq 4096 padding-free pretraining, synthetic packed tokens
  c4  MoEDolomite 24L / 2048d / 16 heads hd 128 / 8 experts top-2 (F 4096 each) / V 50304, seq 2048 pretraining
  c5  Llama-3-8B shape through the `import_from_huggingface` config conversion (32L / 4096d / 32 heads, 8 KV heads, hd 128 /
      F 14336 / V 128256, untied head, no bias), seq 8192 padding-free FINETUNING (ModelWrapperForFinetuning, ragged examples
      with masked prompts), block activation checkpointing

Prints ONE JSON line on rank 0 (contract in the task statement):
  value        whole-job tokens/s with the batches already resident in HBM (device-timed, max over ranks)
  e2e          the same metric through the reference-facing wrapper call (`model({"text": cpu_tensor})` /
               `model({"input_ids": lists, "labels": lists})`) with HOST buffers: H2D of the step's tokens and D2H of the
               loss inside the timed region
  roofline     dominant kernel = the tcgen05 GEMM: algorithmic FLOPs of every GEMM launch / CUDA-event time of that
               launch, measured live inside the timed region, against the measured bf16 peak (MEASURED_PEAKS.json)
  gpu_reference  (N = 1) the reference's own PyTorch / flash-attn path (oracle/gpu_reference.py: cuBLAS nn.Linear,
               flash_attn_varlen_func, eager RMSNorm / RoPE / SwiGLU / CE, autograd, torch AdamW) timed on the same GPU,
               same shape -- the kernel-for-kernel baseline this engine has to beat
  cpu_baseline the oracle (CPU restatement of the reference, torch-eager fp32, eager attention) timed on the host
               cores on a bounded sample (rank 0, N = 1 only)
`--impl reference` times that CPU implementation as its own arm (all host threads), same metric / config.
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

_COMMON = dict(position_embedding_type="rope", activation_function="swiglu", normalization_function="rmsnorm",
               layer_norm_epsilon=1e-5, resid_pdrop=0, embd_pdrop=0, attn_pdrop=0, upcast_logits_for_loss=True)
C2 = dict(model_type="gpt_dolomite", vocab_size=49152, n_positions=4096, n_embd=2560, n_layer=32, n_head=32, n_inner=10240,
          attention_head_type="mha", add_bias=True, eos_token_id=0, **_COMMON)
C4 = dict(model_type="moe_dolomite", vocab_size=50304, n_positions=2048, n_embd=2048, n_layer=24, n_head=16, n_inner=4096,
          num_experts=8, num_experts_per_tok=2, attention_head_type="mha", add_bias=False, eos_token_id=0, **_COMMON)
# Meta-Llama-3-8B config.json as HuggingFace publishes it; converted by hf_models.model_conversion (llama.py:37-74)
LLAMA3_8B_HF = dict(model_type="llama", vocab_size=128256, max_position_embeddings=8192, hidden_size=4096, num_hidden_layers=32,
                    num_attention_heads=32, num_key_value_heads=8, intermediate_size=14336, hidden_act="silu", rms_norm_eps=1e-5,
                    rope_theta=500000.0, attention_bias=False, mlp_bias=False, tie_word_embeddings=False, initializer_range=0.02,
                    bos_token_id=128000, eos_token_id=128001)

WORKLOADS = {
    "c2": dict(seq=4096, mbs=6, kind="pretraining", checkpoint_every=None,
               text="GPTDolomite Granite-3B-code shape (32L/2560d/hd80/F10240/V49152, 3.48B params) bf16 seq4096 padding-free, "
                    "full train step (fwd+bwd+reduce-scatter+clip+AdamW)"),
    "c4": dict(seq=2048, mbs=8, kind="pretraining", checkpoint_every=None,
               text="MoEDolomite 8 experts top-2 (24L/2048d/hd128/F4096 per expert/V50304, 5.4B params) bf16 seq2048 "
                    "padding-free, full train step"),
    "c5": dict(seq=8192, mbs=1, kind="finetuning", checkpoint_every=2,
               text="Llama-3-8B shape via the import_from_huggingface config conversion (32L/4096d/GQA 32:8 hd128/F14336/"
                    "V128256, 8.03B params) bf16 seq8192 padding-free finetuning (4 ragged examples per micro-batch, prompts "
                    "masked), block activation checkpointing, full train step"),
}


def model_config(name: str, layers: int | None = None) -> dict:
    if name == "c2":
        cfg = dict(C2)
    elif name == "c4":
        cfg = dict(C4)
    else:
        from dolomite_engine_b200.hf_models.model_conversion import _import_config

        cfg = _import_config(dict(LLAMA3_8B_HF)).to_dict()
        cfg = {k: v for k, v in cfg.items() if v is not None and k not in ("transformers_version", "architectures")}
        cfg["model_type"] = "gpt_dolomite"
    if layers is not None:
        cfg["n_layer"] = layers
    return cfg


def flops_per_token(cfg: dict, seq: int, checkpointed_fraction: float = 0.0) -> float:
    """reference FLOP model, train_utils.py:197-236 (full SxS attention, no causal discount; recomputed blocks count)"""
    h, f, n, L, v = cfg["n_embd"], cfg["n_inner"], cfg["n_head"], cfg["n_layer"], cfg["vocab_size"]
    k = cfg.get("num_key_value_heads") or n
    mlp = 6 * h * f
    if cfg.get("model_type") == "moe_dolomite":
        mlp = mlp * cfg["num_experts_per_tok"] + 2 * h * cfg["num_experts"]
    fwd = 4 * h * (h * (1 + k / n) + seq) + mlp
    return L * fwd * (3 + checkpointed_fraction) + 6 * h * v


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
# CPU arm: the oracle (reference restatement) on the host cores, bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_reference_tokens_per_s(cfg: dict, seq: int, steps: int = 1, warmup: int = 0, sample_seq: int = 256,
                               budget_s: float = 45.0):
    """One full-depth (all layers) training step of the oracle -- forward, backward, clip and AdamW, fp32 torch-eager with
    eager attention -- on ONE sequence of `sample_seq` tokens.  Nothing is extrapolated over layers; the only reduction of
    the workload is the sequence length (attention's S^2 term is smaller than at the benchmark's S, which favours the CPU).
    If the probe block shows that the full-depth step would blow the time budget the depth is cut and the cut is reported."""
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
    small = {k: v for k, v in cfg.items() if k in O.OracleConfig.__dataclass_fields__}
    small.update(n_positions=max(sample_seq, 16))
    rng = np.random.default_rng(0)
    tokens = rng.integers(0, small["vocab_size"], size=(1, sample_seq + 1), dtype=np.int64)

    def build(n_layer):
        ocfg = O.OracleConfig(**{**small, "n_layer": n_layer})
        params = {k: v.requires_grad_(True) for k, v in O.init_params(ocfg, seed=1).items()}
        opt = torch.optim.AdamW(list(params.values()), lr=1e-5, betas=(0.9, 0.95), eps=1e-10, weight_decay=0.1)
        return ocfg, params, opt

    def one(ocfg, params, opt):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss, _ = O.pretraining_loss(params, ocfg, tokens)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(params.values()), 1.0)
        opt.step()
        return time.perf_counter() - t0

    # probe: one block + head, to size the depth to the budget
    probe = build(1)
    one(*probe)
    t1 = one(*probe)
    import dataclasses

    probe0 = (dataclasses.replace(probe[0], n_layer=0), probe[1], probe[2])  # embedding + LM head only
    one(*probe0)
    t0_ = one(*probe0)
    per_layer = max(t1 - t0_, 1e-3)
    del probe, probe0
    L = cfg["n_layer"]
    runs = max(steps, 1) + max(warmup, 0)
    depth = int(max(1, min(L, (budget_s / runs - t0_) / per_layer)))
    state = build(depth)
    for _ in range(max(warmup, 0)):
        one(*state)
    times = [one(*state) for _ in range(max(steps, 1))]
    t_step = sum(times) / len(times)
    if depth < L:  # depth had to be cut: charge the missing layers at the measured per-layer cost of THIS run
        t_step = t_step + (L - depth) * (t_step - t0_) / depth
    tps = sample_seq / t_step
    return tps, {
        "value": tps, "unit": "tokens/s", "cores": cores, "kind": "port",
        "sample": (f"oracle (torch-eager fp32, eager attention): full train step (fwd+bwd+clip+AdamW) of {depth} of {L} layers + "
                   f"LM head on 1 x {sample_seq} tokens, {t_step:.2f} s/step"
                   + ("" if depth == L else f" after charging the {L - depth} missing layers at the measured per-layer time")
                   + f"; the benchmark's S={seq} attention term is NOT charged (favours the CPU)"),
    }


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    wl = WORKLOADS[args.config]
    cfg = model_config(args.config, args.layers)
    t0 = time.perf_counter()
    tps, cb = cpu_reference_tokens_per_s(cfg, wl["seq"], steps=args.steps, warmup=min(args.warmup, 1))
    tokens_per_step = args.mbs * wl["seq"]
    line = {
        "impl": "reference", "metric": "tokens_per_sec", "value": tps, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tokens_per_step / tps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["text"], "name": args.config, "micro_batch_size": args.mbs, "seq_len": wl["seq"],
                   "n_layer": cfg["n_layer"]},
        "cpu_baseline": cb,
        "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def run_gpu_reference(args) -> None:
    """the reference's PyTorch / flash-attn path on cuda:0 (N = 1): own process, so its memory never coexists with the engine's"""
    import torch

    import oracle.gpu_reference as G

    if int(os.environ.get("RANK", 0)) != 0:
        return
    torch.cuda.set_device(0)
    wl = WORKLOADS[args.config]
    cfg = model_config(args.config, args.layers)
    mbs = args.ref_mbs
    out = {"impl": "gpu_reference", "metric": "tokens_per_sec", "unit": "tokens/s", "n_gpus": 1, "dtype": "bf16",
           "config": {"workload": wl["text"], "name": args.config, "micro_batch_size": mbs, "seq_len": wl["seq"],
                      "n_layer": cfg["n_layer"]},
           "stack": "torch %s: F.linear (cuBLAS), flash_attn_varlen_func, eager RMSNorm/RoPE/SwiGLU/CE, autograd, torch.optim.AdamW"
                    % torch.__version__}
    try:
        import flash_attn

        out["flash_attn"] = flash_attn.__version__
        r = G.time_train_steps(cfg, wl["seq"], mbs, steps=args.steps, warmup=args.warmup, device=torch.device("cuda", 0),
                               docs_per_row=4 if wl["kind"] == "finetuning" else 1)
        out.update(value=r["tokens_per_s"], ms_per_step=r["ms_per_step"], loss=r["loss"], peak_hbm_gb=r["peak_hbm_gb"],
                   steps=args.steps, warmup=args.warmup)
    except Exception as e:  # a baseline that cannot run is reported, never faked
        out.update(value=None, error=f"{type(e).__name__}: {str(e)[:300]}")
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
class _FinetuneFeed:
    """synthetic SFT micro-batches in the padding-free collate format of the reference (data/utils.py:8-92):
    {"input_ids": list[list[int]], "labels": list[list[int]]}; every micro-batch packs `docs` examples whose lengths sum to
    `tokens` (ragged, fixed seed); the first quarter of each example is the prompt (labels -100)."""

    def __init__(self, vocab: int, tokens: int, docs: int, rank: int):
        import numpy as np

        self.rng = np.random.default_rng(4321 + rank)
        self.vocab, self.tokens, self.docs = vocab, tokens, docs

    def __iter__(self):
        return self

    def __next__(self) -> dict:
        cuts = sorted(self.rng.choice(range(64, self.tokens - 64, 8), size=self.docs - 1, replace=False).tolist())
        lens = [b - a for a, b in zip([0] + cuts, cuts + [self.tokens])]
        ids, labels = [], []
        for n in lens:
            x = self.rng.integers(0, self.vocab, size=n).tolist()
            ids.append(x)
            labels.append([-100] * (n // 4) + x[n // 4:])
        return {"input_ids": ids, "labels": labels}


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from dolomite_engine_b200 import _lib
    from dolomite_engine_b200 import kernels as K
    from dolomite_engine_b200.distributed import ShardedDataParallel
    from dolomite_engine_b200.model_wrapper import ModelWrapperForFinetuning, ModelWrapperForPretraining
    from dolomite_engine_b200.optimization import get_optimizer
    from dolomite_engine_b200.pretrain import SyntheticPackedDataset
    from dolomite_engine_b200.train_utils import train_step

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    wl = WORKLOADS[args.config]
    cfg = model_config(args.config, args.layers)
    mbs, seq = args.mbs, wl["seq"]

    # the reference's GPU path first, in its own process (its memory is gone before the engine allocates)
    gpu_ref = None
    if world == 1 and not args.no_gpu_reference:
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "gpu_reference", "--config", args.config, "--steps", "4",
               "--warmup", "2", "--ref-mbs", str(args.ref_mbs)] + (["--layers", str(args.layers)] if args.layers else [])
        try:
            proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
            lines = [l for l in proc.stdout.splitlines() if l.startswith("{")]
            gpu_ref = json.loads(lines[-1]) if lines else {"value": None, "error": (proc.stderr or "no output")[-300:]}
        except Exception as e:
            gpu_ref = {"value": None, "error": f"{type(e).__name__}: {e}"}

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        from dolomite_engine_b200.distributed import configure_comm_ctas

        configure_comm_ctas()
        dist.init_process_group(backend="nccl", rank=rank, world_size=world, device_id=dev)
    finetune = wl["kind"] == "finetuning"
    if finetune:
        wrapper = ModelWrapperForFinetuning(pretrained_config=cfg, device=dev, world_size=world, rank=rank, init_on_device=True)
    else:
        wrapper = ModelWrapperForPretraining(pretrained_config=cfg, micro_batch_size=mbs, sequence_length=seq, device=dev,
                                             world_size=world, rank=rank, init_on_device=True,
                                             reset_attention_mask=args.ragged, reset_position_ids=args.ragged)
    engine = wrapper.model.engine
    ckpt_every = args.checkpoint_every if args.checkpoint_every is not None else wl["checkpoint_every"]
    if ckpt_every:
        engine.checkpoint_every = int(ckpt_every)
    reshard = args.fsdp_mode == "reshard"
    model = ShardedDataParallel(wrapper, dist.group.WORLD if world > 1 else None,
                                communication_dtype={"bf16": torch.bfloat16, "fp32": torch.float32}[args.comm_dtype],
                                reshard_after_forward=reshard)
    opt = get_optimizer("DolomiteFusedAdamW", {"lr": 1e-5, "weight_decay": 0.1, "betas": [0.9, 0.95], "eps": 1e-10}, model)
    if finetune:
        data = _FinetuneFeed(cfg["vocab_size"], mbs * seq, 4 * mbs, rank)
    else:
        data = SyntheticPackedDataset(cfg["vocab_size"], mbs, seq, rank=rank, eos_token_id=cfg["eos_token_id"], ragged=args.ragged)
    tokens_per_step = mbs * seq * world

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- leg 1: batches resident in HBM (4 different batches, used in turn) ----------------
    resident = []
    for _ in range(4):
        b = next(data)
        if finetune:
            from dolomite_engine_b200.hf_models.utils import convert_padding_free_lists_to_tensors

            ids, pos, _, labels, cu, max_seqlen = convert_padding_free_lists_to_tensors(
                input_ids=b["input_ids"], inputs_embeds=None, position_ids=None, token_type_ids=None, labels=b["labels"],
                device=dev)
            resident.append((ids, labels, pos, cu, max_seqlen))
        else:
            ids, labels, pos, cu, max_seqlen = wrapper._stage(b["text"])
            resident.append((ids.clone(), labels.clone(), pos.clone(), cu.clone(), max_seqlen))
    torch.cuda.synchronize()
    turn = [0]

    def resident_step():
        ids, labels, pos, cu, max_seqlen = resident[turn[0] % len(resident)]
        turn[0] += 1
        model.zero_grad()
        model._refresh_parameters_if_needed()
        if finetune:
            loss = wrapper.model(input_ids=ids, position_ids=pos, cu_seqlens=cu, max_seqlen=max_seqlen, labels=labels).loss
        else:
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
        e2e_step()  # includes H2D of the tokens and D2H of loss / grad-norm (.item())
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
        frac_ckpt = 0.0
        if ckpt_every:
            frac_ckpt = len(range(0, cfg["n_layer"], int(ckpt_every))) / cfg["n_layer"]
        fpt = flops_per_token(cfg, seq, frac_ckpt)
        value = tokens_per_step / (ms_resident / 1e3)
        e2e_val = tokens_per_step / (ms_e2e / 1e3)
        gemm_tflops = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        peak = peaks["bf16_tflops_sustained"]
        h2d = wrapper.h2d_bytes_per_step if not finetune else (3 * mbs * seq * 8 + (4 * mbs + 1) * 4)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r02_gemm_traffic_table.json")
        if os.path.exists(tpath):  # written by tools/gemm_traffic_table.py from `ncu --set full` captures of this library
            traffic = json.load(open(tpath))
        line = {
            "metric": "tokens_per_sec", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_resident, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {
                "workload": wl["text"], "name": args.config,
                "micro_batch_size": mbs, "global_batch": mbs * world, "seq_len": seq, "n_layer": cfg["n_layer"],
                "parallelism": f"flat-bucket sharded data parallel x{world}" + (
                    "" if world == 1 else (", stage 3 (parameters resharded after forward, shared gather / gradient buffers)"
                                           if reshard else ", parameters kept gathered between forward and backward")),
                "fsdp_mode": args.fsdp_mode if world > 1 else None,
                "activation_checkpointing": f"every {ckpt_every} block(s)" if ckpt_every else None,
                "packing": "ragged" if (args.ragged or finetune) else "uniform",
                "communication_dtype": args.comm_dtype,
                "gemm_schedule": ("cluster launch control (one cluster per tile, work stealing)" if K.get_option("gemm_dynamic")
                                  else f"static persistent workers, {K.get_option('gemm_sm_margin')} SMs left to NCCL"),
                "attention_cta_order": "heads fastest, longest tiles first" if K.get_option("attn_head_fastest") else "tiles fastest",
                "l2": "working set (parameters + activations per step, tens of GB) far exceeds the 126 MB L2; four resident "
                      "batches are used in turn; no explicit flush",
            },
            "tokens_per_sec_per_gpu": value / world,
            "model_tflops_per_gpu": fpt * value / world / 1e12,
            "model_flops_per_token": fpt,
            "pct_of_bf16_peak_measured_sustained": 100.0 * fpt * value / world / 1e12 / peak,
            "pct_of_bf16_peak_measured_burst": 100.0 * fpt * value / world / 1e12 / peaks["bf16_tflops"],
            "loss": last_loss, "loss_history": [float(x) for x in torch.stack(loss_hist).tolist()],
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "tokens/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8},
            "gpu_launches": launches,
            "peak_hbm_gb": torch.cuda.max_memory_allocated(dev) / 1e9,
            "roofline": {
                "bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05, all nn.Linear fwd/dgrad/wgrad + LM head"
                                             + (", grouped expert GEMMs)" if cfg.get("model_type") == "moe_dolomite" else ")"),
                "achieved": gemm_tflops, "peak": peak, "unit": "TFLOP/s",
                "frac": (gemm_tflops / peak) if gemm_tflops else None,
                # dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures of THIS library,
                # one row per GEMM shape of the step (null until the capture script has run for this commit)
                "traffic": (traffic or {}).get("dominant_launch_dram_bytes"),
                "traffic_algorithmic": (traffic or {}).get("dominant_launch_algorithmic_bytes"),
                "traffic_source": "profiles/r02_gemm_traffic_table.json" if traffic else None,
                "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a long step)",
                "launches_timed": len(gemm_records), "share_of_step": gemm_ms / (ms_resident * args.steps),
            },
        }
        if gpu_ref is not None:
            line["gpu_reference"] = gpu_ref
            if gpu_ref.get("value"):
                line["vs_gpu_reference"] = {"resident": value / gpu_ref["value"], "e2e": e2e_val / gpu_ref["value"],
                                            "note": "tokens/s of this engine / tokens/s of the reference's PyTorch+flash-attn "
                                                    "path on the same GPU (its micro-batch is smaller: eager autograd keeps "
                                                    "more activations)"}
        if world == 1 and not args.no_cpu_baseline:
            try:
                _, cb = cpu_reference_tokens_per_s(cfg, seq, steps=1, warmup=0)
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
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "gpu_reference"])
    ap.add_argument("--config", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--mbs", type=int, default=None,
                    help="sequences per GPU per step; default per workload (c2: 6 x 4096 fills ~146 of 180 GB, measured 41.1 k / "
                         "42.1 k / 42.3 k tokens/s at 4 / 6 / 7)")
    ap.add_argument("--ref-mbs", type=int, default=None, help="micro-batch of the gpu_reference arm (eager autograd keeps more)")
    ap.add_argument("--layers", type=int, default=None, help="override the depth (probes only: a shallower model is not the workload)")
    ap.add_argument("--ragged", action="store_true")
    ap.add_argument("--comm-dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--fsdp-mode", default="reshard", choices=["reshard", "resident"],
                    help="N > 1: reshard = the reference's stage 3 (FULL_SHARD); resident = parameters stay gathered (B200 option)")
    ap.add_argument("--checkpoint-every", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--profile-step", default=None, metavar="JSON",
                    help="diagnostic: write per-kernel device times of one extra (untimed) step to this file")
    args = ap.parse_args()
    wl = WORKLOADS[args.config]
    if args.mbs is None:
        args.mbs = wl["mbs"]
    if args.ref_mbs is None:
        args.ref_mbs = {"c2": 2, "c4": 2, "c5": 1}[args.config]
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "gpu_reference":
        run_gpu_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
