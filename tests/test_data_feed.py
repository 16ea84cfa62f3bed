"""Data feed against fixtures PRODUCED BY THE REFERENCE ITSELF (oracle/pin_data_feed.py ran the
reference's MMapIndexedDatasetBuilder / GPTDataset / helpers.cpp / MegatronBatchSampler in this container): the corpus
files under tests/golden/data_feed/ were written by the reference's builder; expected.npz holds its indices and samples.
Everything here is integer work: bit-exact."""
import filecmp
import os

import numpy as np
import pytest
import torch

from dolomite_engine_b200.data import (
    BlendedDataset,
    GPTDataset,
    MegatronBatchSampler,
    MMapIndexedDataset,
    MMapIndexedDatasetBuilder,
    PackedBatchLoader,
    build_blending_indices,
    build_gpt_datasets,
    build_sample_index,
)
from dolomite_engine_b200.data.gpt_dataset import get_num_epochs, get_split_indices, parse_and_normalize_split

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "data_feed")


@pytest.fixture(scope="module")
def exp():
    return np.load(os.path.join(GOLD, "expected.npz"))


def test_reads_reference_written_store(exp):
    a = MMapIndexedDataset(os.path.join(GOLD, "corpus_a"))
    b = MMapIndexedDataset(os.path.join(GOLD, "corpus_b"))
    assert a.dtype == np.uint16 and b.dtype == np.int32
    assert np.array_equal(a.sequence_lengths, exp["a_sequence_lengths"])
    assert np.array_equal(a.document_indices, exp["a_document_indices"])
    assert np.array_equal(a[5], exp["a_doc5"])
    assert np.array_equal(b.get(3, offset=1, length=4), exp["b_doc3_slice"])
    assert a.sequence_lengths[11] == 0 and a[11].size == 0  # the empty document
    assert len(a[2:5]) == 3 and np.array_equal(a[2:5][1], a[3])


def test_writer_is_byte_identical_to_the_reference_builder(tmp_path):
    for name in ("corpus_a", "corpus_b"):
        src = MMapIndexedDataset(os.path.join(GOLD, name))
        out = str(tmp_path / name)
        w = MMapIndexedDatasetBuilder(out + ".bin", dtype=src.dtype)
        for i in range(len(src)):
            w.add_item(src[i])
            w.end_document()
        w.finalize(out + ".idx")
        assert filecmp.cmp(out + ".bin", os.path.join(GOLD, name + ".bin"), shallow=False)
        assert filecmp.cmp(out + ".idx", os.path.join(GOLD, name + ".idx"), shallow=False)


@pytest.mark.parametrize("ci", range(5))
def test_gpt_dataset_indices_and_samples_match_the_reference(exp, ci):
    lo, hi, num_samples, S, seed, n = (int(x) for x in exp[f"case{ci}_meta"])
    ids = MMapIndexedDataset(os.path.join(GOLD, "corpus_a" if ci < 3 else "corpus_b"))
    ds = GPTDataset(ids, np.arange(lo, hi, dtype=np.int32), num_samples, S, seed)
    assert len(ds) == n
    for name in ("document_index", "sample_index", "shuffle_index"):
        got, want = getattr(ds, name), exp[f"case{ci}_{name}"]
        assert got.dtype == want.dtype and np.array_equal(got, want), name
    take = exp[f"case{ci}_take"]
    got = np.stack([ds[int(i)]["text"] for i in take])
    assert got.dtype == np.int64 and np.array_equal(got, exp[f"case{ci}_samples"])


def test_sample_index_helper_edge_cases(exp):
    sizes, doc_idx = exp["raw_sizes"], exp["raw_doc_idx"]
    tpe = int(sizes.sum())
    for S in (2, 3, 4, 8):
        got32 = build_sample_index(sizes, doc_idx.astype(np.int32), S, 1, tpe)
        got64 = build_sample_index(sizes, doc_idx.astype(np.int64), S, 1, tpe)
        assert got32.dtype == np.int32 and np.array_equal(got32, exp[f"raw_S{S}_i32"])
        assert got64.dtype == np.int64 and np.array_equal(got64, exp[f"raw_S{S}_i64"])
    with pytest.raises(ValueError):  # claims more tokens than the document index holds
        build_sample_index(sizes, doc_idx[:3].astype(np.int32), 4, 1, tpe)


def test_blending_indices_match_the_reference(exp):
    for bi in range(4):
        w = exp[f"blend{bi}_weights"].tolist()
        di, dsi = build_blending_indices(w, exp[f"blend{bi}_dataset_index"].size)
        assert di.dtype == np.int16 and np.array_equal(di, exp[f"blend{bi}_dataset_index"])
        assert np.array_equal(dsi, exp[f"blend{bi}_dataset_sample_index"])


def test_batch_sampler_rank_assignment_and_resume(exp):
    for si in range(3):
        total, consumed, mbs, world, drop_last = (int(x) for x in exp[f"sampler{si}_meta"])
        for rank in range(world):
            rows = list(MegatronBatchSampler(total, consumed, mbs, world, rank, bool(drop_last)))
            lens = exp[f"sampler{si}_rank{rank}_lens"]
            flat = exp[f"sampler{si}_rank{rank}_flat"]
            assert [len(r) for r in rows] == lens.tolist()
            assert [x for r in rows for x in r] == (flat.tolist() if lens.size else [])
    with pytest.raises(AssertionError):
        MegatronBatchSampler(10, 10, 2, 1, 0)


def test_small_helpers():
    assert parse_and_normalize_split("98,1,1") == [0.98, 0.01, 0.01]
    assert parse_and_normalize_split("100") == [1.0, 0.0, 0.0]
    assert get_split_indices([0.9, 0.09, 0.01], 1000) == [0, 900, 990, 1000]
    assert get_split_indices([0.98, 0.01, 0.01], 37)[-1] == 37
    for tpe, S, n in [(100, 8, 5), (100, 8, 12), (100, 8, 13), (7, 3, 50), (1000, 999, 1)]:
        e = get_num_epochs(tpe, S, n)
        assert (e * tpe - 1) // S >= n and (e == 1 or ((e - 1) * tpe - 1) // S < n)


def test_loader_emits_wrapper_batches_and_ranks_partition_the_global_batch():
    S, mbs, world = 16, 3, 2
    ids = MMapIndexedDataset(os.path.join(GOLD, "corpus_a"))
    ds = GPTDataset(ids, np.arange(0, 37, dtype=np.int32), 60, S, 1234)
    per_rank = []
    for rank in range(world):
        ld = PackedBatchLoader(ds, MegatronBatchSampler(len(ds), 6, mbs, world, rank), S, pin=False)
        batches = [b["text"] for b in ld]
        per_rank.append(batches)
        assert all(b.dtype == torch.int64 and tuple(b.shape) == (mbs, S + 1) for b in batches)
        assert ld.consumed_samples == 6 + len(batches) * mbs * world  # resume point (consumed_samples of the reference)
    n_steps = (len(ds) - 6) // (mbs * world)
    assert len(per_rank[0]) == len(per_rank[1]) == n_steps
    for step in range(n_steps):
        for rank in range(world):
            for r in range(mbs):
                want = ds[6 + step * mbs * world + rank * mbs + r]["text"]
                assert np.array_equal(per_rank[rank][step][r].numpy(), want)
    # consecutive samples of the UNSHUFFLED stream overlap by exactly one token (labels of the last position)
    inv = np.argsort(ds.shuffle_index)
    a, b = ds[int(inv[0])]["text"], ds[int(inv[1])]["text"]
    assert a[-1] == b[0]


def test_blended_dataset_and_builder(tmp_path):
    S = 8
    a = MMapIndexedDataset(os.path.join(GOLD, "corpus_a"))
    b = MMapIndexedDataset(os.path.join(GOLD, "corpus_b"))
    da = GPTDataset(a, np.arange(0, 37, dtype=np.int32), 80, S, 1)
    db = GPTDataset(b, np.arange(0, 23, dtype=np.int32), 80, S, 1)
    bl = BlendedDataset([da, db], [0.25, 0.75], 64)
    counts = np.bincount(bl.dataset_index, minlength=2)
    assert abs(counts[0] - 16) <= 1 and counts.sum() == 64
    for i in (0, 1, 17, 63):
        src, j = bl.locate(i)
        assert np.array_equal(bl[i]["text"], src[j]["text"])
    with pytest.raises(IndexError):
        bl[64]
    ld = PackedBatchLoader(bl, MegatronBatchSampler(len(bl), 0, 4, 1, 0), S, pin=False)
    first = next(iter(ld))["text"]
    assert np.array_equal(first.numpy(), np.stack([bl[i]["text"] for i in range(4)]))
    train, val, test = build_gpt_datasets(["1", os.path.join(GOLD, "corpus_a"), "3", os.path.join(GOLD, "corpus_b")],
                                          "90,10,0", (40, 8, 0), S, 1234)
    # the blend's length is the sum of the per-store requests incl. the 0.5 % margin: ceil(40 * .25 * 1.005) + ceil(40 * .75 * 1.005)
    assert isinstance(train, BlendedDataset) and len(train) == 11 + 31 and len(val) == 3 + 7 and test is None
    t1, v1, _ = build_gpt_datasets(os.path.join(GOLD, "corpus_a"), "80,20,0", (30, 5, 0), S, 1234)
    assert isinstance(t1, GPTDataset) and t1.indexed_indices[-1] + 1 == v1.indexed_indices[0]


def test_yaml_megatron_dataset_feeds_the_training_loop_contract():
    """class_name: MegatronDataset in the YAML -> pretrain.make_dataloader -> {"text": int64[mbs, S+1]} (pretraining.py:89)"""
    from dolomite_engine_b200.arguments import get_args_from_dict, load_yaml
    from dolomite_engine_b200.pretrain import make_dataloader

    d = load_yaml(os.path.join(os.path.dirname(HERE), "configs", "c1_tiny.yml"))
    d["datasets"] = [dict(class_name="MegatronDataset", data_name="Megatron",
                          class_args=dict(data_path=[os.path.join(GOLD, "corpus_a")], split="100,0,0", sequence_length=16,
                                          eval_steps=2, seed=7))]
    d["training_parameters"].update(num_training_steps=4, micro_batch_size=2, gradient_accumulation_steps=1, eval_interval=2)
    args = get_args_from_dict(d)
    seen = []
    for rank in range(2):
        it = make_dataloader(args, None, rank, world=2)
        seen.append([next(it)["text"] for _ in range(4)])
        assert all(tuple(t.shape) == (2, 17) and t.dtype == torch.int64 for t in seen[-1])
    assert not torch.equal(seen[0][0], seen[1][0])  # ranks read different rows of the same global batch
    # resuming at consumed_samples = 2 steps * (mbs * world) reproduces step 2 onwards
    it = make_dataloader(args, None, 0, world=2, consumed_samples=2 * 2 * 2)
    assert torch.equal(next(it)["text"], seen[0][2])


def test_yaml_data_cache_path_stores_the_indices_and_feeds_the_same_batches(tmp_path):
    """class_args.data_cache_path (data/megatron/__init__.py:85): the indices are stored under the reference's file names and a second
    start maps them instead of rebuilding; the token stream does not change"""
    from dolomite_engine_b200.arguments import get_args_from_dict, load_yaml
    from dolomite_engine_b200.pretrain import make_dataloader

    d = load_yaml(os.path.join(os.path.dirname(HERE), "configs", "c1_tiny.yml"))
    ca = dict(data_path=[os.path.join(GOLD, "corpus_a")], split="100,0,0", sequence_length=16, eval_steps=2, seed=7)
    d["training_parameters"].update(num_training_steps=4, micro_batch_size=2, gradient_accumulation_steps=1, eval_interval=2)
    batches = []
    for extra in ({}, {"data_cache_path": str(tmp_path / "idx")}, {"data_cache_path": str(tmp_path / "idx")}):
        d["datasets"] = [dict(class_name="MegatronDataset", data_name="Megatron", class_args={**ca, **extra})]
        loader = make_dataloader(get_args_from_dict(d), None, 0, world=1)
        batches.append([next(loader)["text"].clone() for _ in range(3)])
    assert all(torch.equal(a, b) and torch.equal(a, c) for a, b, c in zip(*batches))
    names = sorted(os.listdir(tmp_path / "idx"))
    assert len(names) == 4 and all("-GPTDataset-" in n for n in names)
    assert not os.path.exists(os.path.join(GOLD, "corpus_a", "cache"))  # nothing is written next to the data without being asked


def test_validation_split_loader_and_evaluate_contract():
    """pretrain.evaluate: mean of the per-batch losses over eval_steps batches of the validation split, model back in train mode"""
    from dolomite_engine_b200.arguments import get_args_from_dict, load_yaml
    from dolomite_engine_b200.pretrain import evaluate, make_megatron_val_dataloader

    d = load_yaml(os.path.join(os.path.dirname(HERE), "configs", "c1_tiny.yml"))
    d["datasets"] = [dict(class_name="MegatronDataset", data_name="Megatron",
                          class_args=dict(data_path=[os.path.join(GOLD, "corpus_a")], split="70,30,0", sequence_length=16,
                                          eval_steps=3, seed=7))]
    d["training_parameters"].update(num_training_steps=4, micro_batch_size=2, gradient_accumulation_steps=1, eval_interval=2,
                                    eval_during_training=True)
    args = get_args_from_dict(d)
    factory = make_megatron_val_dataloader(args, rank=0, world=1)
    assert factory is not None
    first = [b["text"].clone() for _, b in zip(range(3), factory())]
    again = [b["text"].clone() for _, b in zip(range(3), factory())]
    assert all(torch.equal(a, b) for a, b in zip(first, again))  # every evaluation restarts at sample 0

    class Fake(torch.nn.Module):
        def forward(self, batch):
            return batch["text"].float().mean()

    m = Fake()
    v = evaluate(factory, m, 3, 1)
    assert m.training and abs(v - float(torch.stack([t.float().mean() for t in first]).mean())) < 1e-4
    d["datasets"][0]["class_args"]["split"] = "100,0,0"
    assert make_megatron_val_dataloader(get_args_from_dict(d), 0, 1) is None


def test_preprocess_jsonl_to_token_stores(tmp_path):
    """data/preprocess.py (reference tool tools/megatron_dataset/preprocess_data.py): one store per JSON key, one document
    per line, eod appended, empty documents skipped, uint16 for small vocabularies"""
    import json

    from dolomite_engine_b200.data import MMapIndexedDataset
    from dolomite_engine_b200.data import preprocess as P

    src = tmp_path / "c.jsonl"
    rows = [{"text": "hello world", "code": "ab"}, {"text": "", "code": "xyz"}, {"text": "q", "code": ""}]
    src.write_text("\n".join(json.dumps(r) for r in rows) + "\n\n")
    P.configure(lambda s: [ord(c) for c in s], ["text", "code"], eod=1)
    counts = P.write_stores(map(P.encode_line, P._lines(str(src))), str(tmp_path / "out"), ["text", "code"], vocab_size=300)
    assert counts == {"text": 2, "code": 2}
    text = MMapIndexedDataset(str(tmp_path / "out_text"))
    assert text.dtype == np.uint16 and list(text.sequence_lengths) == [12, 2]
    assert list(text[0]) == [ord(c) for c in "hello world"] + [1] and list(text[1]) == [ord("q"), 1]
    assert list(text.document_indices) == [0, 1, 2]
    code = MMapIndexedDataset(str(tmp_path / "out_code"))
    assert [list(code[i]) for i in range(2)] == [[97, 98, 1], [120, 121, 122, 1]]
    wide = P.write_stores([{"text": [70000, 5]}], str(tmp_path / "w"), ["text"], vocab_size=128000)
    assert wide == {"text": 1} and MMapIndexedDataset(str(tmp_path / "w_text")).dtype == np.int32


def test_builder_roundtrip_many_documents(tmp_path):
    """mirrors the reference's tests/data/megatron_data_test.py::test_megatron_dataset_builder"""
    from dolomite_engine_b200.data import MMapIndexedDataset, MMapIndexedDatasetBuilder, get_bin_path, get_idx_path

    prefix = str(tmp_path / "file")
    b = MMapIndexedDatasetBuilder(get_bin_path(prefix))
    for _ in range(1000):
        b.add_item(np.array([1, 2]))
        b.end_document()
    b.finalize(get_idx_path(prefix))
    ds = MMapIndexedDataset(prefix)
    assert len(ds) == 1000 and all((ds[i] == [1, 2]).all() for i in range(1000))


def test_merge_matches_reference_add_index(tmp_path):
    """data/merge.py + builder.add_index: merged files are byte-identical to what the reference's builder wrote
    (oracle/pin_merge.py -> tests/golden/merge_expected.json); mirrors megatron_data_test.py::test_megatron_dataset_merge"""
    import hashlib
    import json

    from dolomite_engine_b200.data import MMapIndexedDataset
    from dolomite_engine_b200.data.merge import merge, prefixes_in

    golden = os.path.join(HERE, "golden")
    want = json.load(open(os.path.join(golden, "merge_expected.json")))
    for name, w in want.items():
        out = str(tmp_path / name)
        n = merge([os.path.join(golden, p) for p in w["parts"]], out)
        assert n == w["sequences"]
        for ext in ("bin", "idx"):
            assert hashlib.sha256(open(f"{out}.{ext}", "rb").read()).hexdigest() == w[ext], (name, ext)
        ds = MMapIndexedDataset(out)
        assert ds.document_indices.shape[0] == w["documents"]
    a = MMapIndexedDataset(os.path.join(golden, "data_feed", "corpus_a"))
    merged = MMapIndexedDataset(str(tmp_path / "a_fim_a"))
    assert all(np.array_equal(merged[i], a[i]) for i in range(len(a)))  # first part unchanged
    assert all(np.array_equal(merged[len(merged) - len(a) + i], a[i]) for i in range(len(a)))  # last part appended intact
    with pytest.raises(ValueError):  # token widths must agree
        merge([os.path.join(golden, "data_feed", "corpus_a"), os.path.join(golden, "data_feed", "corpus_b")], str(tmp_path / "bad"))
    assert [os.path.basename(p) for p in prefixes_in(os.path.join(golden, "data_feed"))] == ["corpus_a", "corpus_b"]


def test_split_arithmetic_matches_reference():
    """`split: "969,30,1"` -> normalised weights -> document index bounds, as evaluated by the reference's own functions
    (oracle/pin_split_logic.py -> tests/golden/split_logic.json)"""
    import json

    for case in json.load(open(os.path.join(HERE, "golden", "split_logic.json"))):
        vec = parse_and_normalize_split(case["split"])
        assert list(vec) == case["vector"], case["split"]
        for n, bounds in case["bounds"].items():
            assert list(get_split_indices(vec, int(n))) == bounds, (case["split"], n)


# ------------------------------------------------------------------------------------------------
# index cache (gpt_dataset.py:241-400 of the reference): same description, same MD5, same file names, interchangeable files
# ------------------------------------------------------------------------------------------------
_CACHE_CASE = dict(num_samples=230, sequence_length=16, random_seed=77, index_split="train", split="90,10,0")


def _cache_case_dataset(**kw):
    ids = MMapIndexedDataset("corpus_a")  # relative prefix: the description (hence the MD5) must not depend on the checkout path
    c = _CACHE_CASE
    return GPTDataset(ids, np.arange(0, 33, dtype=np.int32), c["num_samples"], c["sequence_length"], c["random_seed"],
                      index_split=c["index_split"], split=c["split"], **kw)


def test_index_cache_written_by_the_reference_is_found_and_used(exp, monkeypatch):
    monkeypatch.chdir(GOLD)
    ds = _cache_case_dataset(path_to_cache="ref_index_cache", cache="load")
    assert ds.unique_description == str(exp["cache_description"]) and ds.unique_description_hash == str(exp["cache_hash"])
    assert ds.cache_hit and isinstance(ds.sample_index, np.memmap) and isinstance(ds.shuffle_index, np.memmap)
    assert sorted(os.path.basename(p) for p in ds.cache_paths().values()) == list(exp["cache_files"])
    assert len(ds) == int(exp["cache_len"]) and ds.num_epochs >= 2
    got = np.stack([ds[i]["text"] for i in range(0, len(ds), 7)])
    assert np.array_equal(got, exp["cache_samples"])
    # the in-memory build produces exactly what the reference stored (values and dtypes)
    mem = _cache_case_dataset()
    assert not mem.cache_hit
    for a, b in ((mem.document_index, ds.document_index), (mem.sample_index, ds.sample_index), (mem.shuffle_index, ds.shuffle_index)):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    # the validation split of the same configuration hashes like the reference's
    ids = MMapIndexedDataset("corpus_a")
    dv = GPTDataset(ids, np.arange(33, 37, dtype=np.int32), 9, 16, 77, index_split="valid", split="90,10,0",
                    path_to_cache="ref_index_cache", cache="load")
    assert dv.unique_description_hash == str(exp["cache_hash_valid"]) and dv.cache_hit


def test_index_cache_build_store_reload_and_unwritable_directory(exp, monkeypatch, tmp_path):
    monkeypatch.chdir(GOLD)
    cache = str(tmp_path / "cache")
    first = _cache_case_dataset(path_to_cache=cache, cache="build")
    assert not first.cache_hit and isinstance(first.sample_index, np.memmap)  # built, stored, mapped back
    assert sorted(os.listdir(cache)) == list(exp["cache_files"])  # the reference's file names; no temporary file left behind
    for name in exp["cache_files"]:
        ours, ref = os.path.join(cache, str(name)), os.path.join("ref_index_cache", str(name))
        if str(name).endswith(".npy"):
            a, b = np.load(ours), np.load(ref)
            assert a.dtype == b.dtype and np.array_equal(a, b)
        else:
            assert open(ours).read() == open(ref).read()
    again = _cache_case_dataset(path_to_cache=cache, cache="build")
    assert again.cache_hit and np.array_equal(again[5]["text"], first[5]["text"])
    # "load" never writes: a miss builds in memory
    empty = str(tmp_path / "empty")
    miss = _cache_case_dataset(path_to_cache=empty, cache="load")
    assert not miss.cache_hit and not os.path.exists(empty) and np.array_equal(miss[5]["text"], first[5]["text"])
    # a cache directory that cannot be created degrades to the in-memory indices (with a warning), it does not abort the run
    blocker = tmp_path / "file"
    blocker.write_text("x")
    with pytest.warns(UserWarning, match="not cached"):
        bad = _cache_case_dataset(path_to_cache=str(blocker / "sub"), cache="build")
    assert np.array_equal(bad[5]["text"], first[5]["text"])
    with pytest.raises(ValueError):
        _cache_case_dataset(cache="sometimes")


def test_build_gpt_datasets_stores_train_and_validation_indices(tmp_path):
    cache = str(tmp_path / "c")
    prefix = os.path.join(GOLD, "corpus_a")
    t, v, _ = build_gpt_datasets(prefix, "80,20,0", (30, 5, 0), 16, 1234, data_cache_path=cache, cache="build")
    files = sorted(os.listdir(cache))
    assert len(files) == 8 and {f.split("-")[0] for f in files} == {t.unique_description_hash, v.unique_description_hash}
    t2, v2, _ = build_gpt_datasets(prefix, "80,20,0", (30, 5, 0), 16, 1234, data_cache_path=cache, cache="load")
    assert t2.cache_hit and v2.cache_hit and np.array_equal(t2[3]["text"], t[3]["text"]) and np.array_equal(v2[1]["text"], v[1]["text"])
    # default directory of the reference: <path_prefix>/cache/GPTDataset_indices (nothing is written there in "load" mode)
    t3, _, _ = build_gpt_datasets(prefix, "80,20,0", (30, 5, 0), 16, 1234, cache="load")
    assert os.path.dirname(t3.cache_paths()["sample_index.npy"]) == os.path.join(prefix, "cache", "GPTDataset_indices")
    assert not t3.cache_hit and not os.path.exists(os.path.join(prefix, "cache"))


def _cache_worker(rank: int, world: int, port: int, cache: str, q):
    import sys

    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        t, _, _ = build_gpt_datasets(os.path.join(GOLD, "corpus_b"), "100,0,0", (57, 0, 0), 32, 1234, data_cache_path=cache, cache="build")
        q.put((rank, bool(t.cache_hit), isinstance(t.sample_index, np.memmap), t[4]["text"].tolist()))
    except Exception:  # noqa
        import traceback

        q.put((rank, traceback.format_exc(), None, None))
    finally:
        dist.destroy_process_group()


def test_rank_0_builds_the_index_cache_and_the_other_ranks_map_it(tmp_path, exp):
    """blended_megatron_dataset_builder.py:330-366: caching ranks build first, a barrier, then everybody else loads"""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 77) % 2000
    procs = [ctx.Process(target=_cache_worker, args=(r, 2, port, str(tmp_path / "shared"), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    (r0, hit0, map0, s0), (r1, hit1, map1, s1) = res
    assert isinstance(hit0, bool), hit0
    assert isinstance(hit1, bool), hit1
    assert (r0, r1) == (0, 1) and not hit0 and hit1 and map0 and map1 and s0 == s1
    assert len(os.listdir(tmp_path / "shared")) == 4


def _check_split_against_the_reference_builder(exp, tag, i, ds):
    n = int(exp[f"builder_{tag}_lens"][i])
    if n < 0:
        assert ds is None
        return
    assert len(ds) == n
    want = exp[f"builder_{tag}_split{i}_samples"]
    assert np.array_equal(np.stack([ds[j]["text"] for j in range(want.shape[0])]), want)
    if f"builder_{tag}_split{i}_dataset_index" in exp.files:
        assert np.array_equal(ds.dataset_index, exp[f"builder_{tag}_split{i}_dataset_index"])
        assert np.array_equal(ds.dataset_sample_index, exp[f"builder_{tag}_split{i}_dataset_sample_index"])


def test_builder_options_match_the_reference_builder(exp):
    """what BlendedMegatronDatasetBuilder.build() itself returned in this container (oracle/pin_data_feed.py): option 2 = weighted
    blend cut by `split`; option 3 = one blend per split (class_args train_data_path / val_data_path / test_data_path)"""
    pa, pb = os.path.join(GOLD, "corpus_a"), os.path.join(GOLD, "corpus_b")
    got = build_gpt_datasets(["1", pa, "3", pb], "90,10,0", (40, 8, 0), 8, 1234)
    for i in range(3):
        _check_split_against_the_reference_builder(exp, "opt2", i, got[i])
    got = build_gpt_datasets(None, None, (50, 6, 0), 8, 5, blend_per_split=[["2", pa, "1", pb], [pb], None])
    for i in range(3):
        _check_split_against_the_reference_builder(exp, "opt3", i, got[i])
    assert got[1].unique_description.count('"split": null') == 1 and '"index_split": "valid"' in got[1].unique_description
    with pytest.raises(AssertionError):
        build_gpt_datasets(pa, "100,0,0", (5, 0, 0), 8, 5, blend_per_split=[[pb], None, None])
    with pytest.raises(AssertionError):
        build_gpt_datasets(pa, None, (5, 0, 0), 8, 5)


def test_yaml_train_and_val_data_paths_feed_their_own_stores(exp):
    """class_args train_data_path / val_data_path (option 3): the training loader reads the train blend, the validation loader the
    validation store; the first rows are the ones the reference's builder produced"""
    from dolomite_engine_b200.arguments import get_args_from_dict, load_yaml
    from dolomite_engine_b200.pretrain import make_dataloader, make_megatron_val_dataloader

    pa, pb = os.path.join(GOLD, "corpus_a"), os.path.join(GOLD, "corpus_b")
    d = load_yaml(os.path.join(os.path.dirname(HERE), "configs", "c1_tiny.yml"))
    d["datasets"] = [dict(class_name="MegatronDataset", data_name="Megatron",
                          class_args=dict(train_data_path=["2", pa, "1", pb], val_data_path=[pb], sequence_length=8, eval_steps=3, seed=5))]
    # train request = steps * mbs * accumulation * world = 50 samples like the pinned builder case; the validation store is read through a
    # single GPTDataset, which holds whole epochs whatever the request (same first rows as the pinned case)
    d["training_parameters"].update(num_training_steps=25, micro_batch_size=2, gradient_accumulation_steps=1, eval_interval=25,
                                    eval_during_training=True)
    args = get_args_from_dict(d)
    it = make_dataloader(args, None, 0, world=1)
    rows = torch.cat([next(it)["text"] for _ in range(3)]).numpy()
    assert np.array_equal(rows, exp["builder_opt3_split0_samples"][:6])
    val = make_megatron_val_dataloader(args, rank=0, world=1)
    vrows = torch.cat([b["text"] for _, b in zip(range(2), val())]).numpy()
    assert np.array_equal(vrows, exp["builder_opt3_split1_samples"][:4])
    d["datasets"][0]["class_args"]["data_path"] = [pa]
    with pytest.raises(ValueError):
        make_dataloader(get_args_from_dict(d), None, 0, world=1)
