"""CPU, world_size 2, gloo: the host-side sharding contract of the flat-bucket data-parallel runtime
(dolomite_engine_b200/distributed.py) -- the slices every rank owns tile the padded unit, gathering the shards
reproduces every named parameter, reduce-scatter(AVG) of rank-local gradients leaves each rank the mean of its slice,
and the global gradient norm is sqrt(all_reduce(sum of shard sum-of-squares)).  The NCCL/stream side runs on the GPU
box (bench.py --gpus N, tests marked gpu)."""

import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank: int, world: int, port: int, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dolomite_engine_b200.engine import FlatUnit, _block_specs, _root_specs
        from dolomite_engine_b200.hf_models import GPTDolomiteConfig

        cfg = GPTDolomiteConfig(n_embd=64, n_head=4, n_layer=2, n_inner=96, vocab_size=264, attention_head_type="gqa",
                                num_key_value_heads=2, add_bias=True, activation_function="swiglu",
                                position_embedding_type="rope", normalization_function="rmsnorm")
        units = [FlatUnit("root", _root_specs(cfg), world, rank)] + [
            FlatUnit(f"h.{i}", _block_specs(cfg, i), world, rank) for i in range(cfg.n_layer)]
        g = torch.Generator().manual_seed(42)  # same seed on every rank -> same full parameters
        total_sq = torch.zeros(1)
        for u in units:
            u.allocate("cpu")
            full = u.init_full(g)
            u.full_master_from(full)
            # (1) shards tile the padded buffer
            assert u.shard_numel * world == u.padded and u.master.numel() == u.shard_numel
            # (2) all-gather of the fp32 shards reproduces every named parameter
            parts = [torch.empty(u.shard_numel) for _ in range(world)]
            dist.all_gather(parts, u.master.data)
            gathered = torch.cat(parts)
            assert torch.equal(gathered, full)
            for s in u.specs:
                assert torch.equal(gathered[s.offset : s.offset + s.numel].view(s.shape), full[s.offset : s.offset + s.numel].view(s.shape))
            # (3) reduce-scatter(AVG) semantics on rank-local full gradients
            local = torch.full((u.padded,), float(rank + 1)) + torch.arange(u.padded) * 1e-3
            summed = local.clone()
            dist.all_reduce(summed)
            mean_slice = (summed / world)[rank * u.shard_numel : (rank + 1) * u.shard_numel]
            expect = (torch.full((u.padded,), (1 + world) / 2) + torch.arange(u.padded) * 1e-3)[rank * u.shard_numel : (rank + 1) * u.shard_numel]
            assert torch.allclose(mean_slice, expect)
            u.master.grad.copy_(mean_slice)
            total_sq += u.master.grad.double().pow(2).sum().float()
        # (4) global grad norm from shard-local sums of squares
        dist.all_reduce(total_sq)
        ref = torch.cat([(torch.full((u.padded,), (1 + world) / 2) + torch.arange(u.padded) * 1e-3) for u in units]).double().norm()
        assert abs(total_sq.sqrt().item() - ref.item()) / ref.item() < 1e-5
        q.put((rank, "ok"))
    except Exception as e:  # noqa
        import traceback

        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_world_size_2_sharding_contract():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status in results:
        assert status == "ok", f"rank {rank}: {status}"
