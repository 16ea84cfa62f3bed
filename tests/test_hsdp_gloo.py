"""CPU, world_size 4, gloo: the 2-D (replicate x shard) data-parallel topology of `zero_topology`
(reference arguments.py:283-298, utils/parallel.py:59-68, :255-266): group layout, and that
reduce-scatter(AVG) inside the shard group followed by all-reduce(AVG) between the owners of a slice leaves every rank
the GLOBAL mean of its slice, with the gradient norm summed over ONE shard group.  The NCCL / stream side is the same
code path as plain sharding plus one collective (distributed._Comm.post_backward_unit)."""

import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dolomite_engine_b200.distributed import hybrid_layout


def test_hybrid_layout():
    s, r = hybrid_layout(8, None, None)
    assert s == [list(range(8))] and r == [[i] for i in range(8)]
    s, r = hybrid_layout(8, 4, 2)
    assert s == [[0, 1, 2, 3], [4, 5, 6, 7]]  # consecutive ranks (one NVSwitch domain) shard
    assert r == [[0, 4], [1, 5], [2, 6], [3, 7]]
    s, r = hybrid_layout(6, 2, 3)
    assert s == [[0, 1], [2, 3], [4, 5]] and r == [[0, 2, 4], [1, 3, 5]]
    # every rank is in exactly one group of each kind
    for groups in (s, r):
        assert sorted(x for g in groups for x in g) == list(range(6))
    with pytest.raises(ValueError):
        hybrid_layout(8, 3, 2)
    with pytest.raises(ValueError):
        hybrid_layout(8, 4, None)


def test_arguments_accept_zero_topology():
    from dolomite_engine_b200.arguments import DistributedArgs

    d = DistributedArgs(zero_topology={"data_parallel_replication_world_size": 2, "data_parallel_sharding_world_size": 4})
    assert d.zero_topology.data_parallel_sharding_world_size == 4
    with pytest.raises(Exception):
        DistributedArgs(zero_topology={"data_parallel_replication_world_size": 2})


def _worker(rank: int, world: int, port: int, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dolomite_engine_b200.distributed import build_data_parallel_groups
        from dolomite_engine_b200.engine import FlatUnit, _block_specs
        from dolomite_engine_b200.hf_models import GPTDolomiteConfig

        S, R = 2, 2
        shard_group, replicate_group, shard_world, shard_rank = build_data_parallel_groups(S, R)
        assert shard_world == S and shard_rank == rank % S
        assert dist.get_world_size(shard_group) == S and dist.get_world_size(replicate_group) == R
        assert dist.get_rank(shard_group) == shard_rank and dist.get_rank(replicate_group) == rank // S

        cfg = GPTDolomiteConfig(n_embd=64, n_head=4, n_layer=1, n_inner=96, vocab_size=264, attention_head_type="gqa",
                                num_key_value_heads=2, add_bias=True, activation_function="swiglu",
                                position_embedding_type="rope", normalization_function="rmsnorm")
        u = FlatUnit("h.0", _block_specs(cfg, 0), shard_world, shard_rank)
        u.allocate("cpu")
        g = torch.Generator().manual_seed(42)
        full = u.init_full(g)
        u.full_master_from(full)
        # replicas hold identical shards
        other = [torch.empty_like(u.master.data) for _ in range(R)]
        dist.all_gather(other, u.master.data, group=replicate_group)
        assert all(torch.equal(o, u.master.data) for o in other)
        # the shard group alone reproduces the full parameters
        parts = [torch.empty(u.shard_numel) for _ in range(S)]
        dist.all_gather(parts, u.master.data, group=shard_group)
        assert torch.equal(torch.cat(parts), full)

        # gradient path: rank-local full gradients -> RS(AVG) in the shard group -> AR(AVG) across replicas
        gl = torch.Generator().manual_seed(100 + rank)
        local = torch.randn(u.padded, generator=gl)
        chunks = list(local.chunk(S))
        out = torch.empty(u.shard_numel)
        dist.reduce_scatter(out, chunks, op=dist.ReduceOp.SUM, group=shard_group)  # gloo has no AVG
        out /= S
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=replicate_group)
        out /= R
        every = [torch.randn(u.padded, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
        mean = torch.stack(every).mean(0)
        lo = shard_rank * u.shard_numel
        assert torch.allclose(out, mean[lo : lo + u.shard_numel], atol=1e-6)
        # gradient norm: sum of squares over ONE shard group (replicas are identical)
        sq = out.double().pow(2).sum().float().reshape(1)
        dist.all_reduce(sq, group=shard_group)
        assert abs(sq.sqrt().item() - mean.double().norm().item()) < 1e-4
        q.put((rank, "ok"))
    except Exception:  # noqa
        import traceback

        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_world_size_4_replicate_x_shard():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 4, port, q)) for r in range(4)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status in results:
        assert status == "ok", f"rank {rank}: {status}"
