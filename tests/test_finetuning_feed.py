"""SFT feed (decoder-only) against fixtures produced by the reference's `collate_fn` (data/utils.py:8-92) and
`BaseDataset.get_input_output_token_ids` (data/base.py:84-121) -- tests/golden/finetuning_feed.json, written by
oracle/pin_finetuning_feed.py in the build container.  Integer work: exact."""
import json
import os

import pytest

import torch

from dolomite_engine_b200.data.finetuning import JSONLinesSFTDataset, batches, build_example, collate

HERE = os.path.dirname(os.path.abspath(__file__))


def toy_tokenize(text):
    return [3 + (sum(map(ord, w)) % 97) for w in text.split()]


def _fx():
    return json.load(open(os.path.join(HERE, "golden", "finetuning_feed.json")))


def test_examples_match_reference_tokenisation_and_truncation():
    fx = _fx()
    for group in fx["examples"]:
        got = [build_example(toy_tokenize, 2, i, o, group["max_input_tokens"], group["max_output_tokens"]) for i, o in fx["raw"]]
        assert got == group["examples"]


def test_collate_matches_reference_padding_free_and_left_padded():
    fx = _fx()
    by_limits = {(g["max_input_tokens"], g["max_output_tokens"]): g["examples"] for g in fx["examples"]}
    for case in fx["collate"]:
        exs = by_limits[(case["max_input_tokens"], case["max_output_tokens"])]
        got = collate(exs, 2, case["padding_free"], case["loss_mask"])
        want = case["result"]
        assert set(got) == set(want)
        for k, v in want.items():
            g = got[k].tolist() if torch.is_tensor(got[k]) else got[k]
            assert g == v, (case["padding_free"], case["loss_mask"], k)
        if not case["padding_free"]:
            assert got["input_ids"].dtype == torch.long and got["attention_mask"].sum(1).tolist() == [len(e["input"]) for e in exs]


def test_jsonl_dataset_and_rank_sharded_batches(tmp_path):
    rows = [{"input": f"question number {i} please", "output": "answer " * (1 + i % 3)} for i in range(11)]
    with open(tmp_path / "train.jsonl", "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
    ds = JSONLinesSFTDataset(str(tmp_path), toy_tokenize, 2, input_format="Q: __input__\nA:", max_output_tokens=3)
    assert len(ds) == 11 and all(len(e["output"]) <= 3 and e["output"][-1] == 2 for e in ds.examples)
    assert ds[0]["input"][: len(toy_tokenize("Q: question number 0 please A:"))] == toy_tokenize("Q: question number 0 please\nA:")
    seen = []
    for rank in range(2):
        it = batches(ds, 2, 2, True, rank=rank, world_size=2, seed=5, infinite=False)
        got = list(it)
        assert all(len(b["input_ids"]) == 2 and len(b["labels"]) == 2 for b in got)
        seen.append([tuple(x) for b in got for x in b["input_ids"]])
    assert not set(seen[0]) & set(seen[1])  # ranks see disjoint examples
    padded = next(batches(ds, 3, 2, False, seed=5))
    assert padded["input_ids"].shape == padded["attention_mask"].shape == padded["labels"].shape
    assert (padded["labels"][padded["attention_mask"] == 0] == -100).all()


def test_finetune_yaml_to_batches(tmp_path):
    """finetune.make_sft_dataloader: YAML dataset entry (formats, token limits, loss mask) -> collated micro-batches"""
    from dolomite_engine_b200.arguments import get_args_from_dict, load_yaml
    from dolomite_engine_b200.finetune import make_sft_dataloader

    with open(tmp_path / "d.jsonl", "w") as f:
        for i in range(9):
            f.write(json.dumps({"input": f"prompt {i} a b c d e", "output": f"reply {i} x y"}) + "\n")
    d = load_yaml(os.path.join(os.path.dirname(HERE), "configs", "c1_tiny.yml"))
    d["datasets"] = [dict(class_name="JSONLinesDataset", data_name="sft", class_args=dict(data_path=str(tmp_path)),
                          input_format="Q: __input__ A:", output_format="__output__", max_input_tokens=5, max_output_tokens=3)]
    d["tuning_args"] = dict(tuning_method="full_finetuning")
    d["model_args"]["use_padding_free_transformer"] = False
    d["training_parameters"].update(micro_batch_size=4, loss_mask="output_only")
    args = get_args_from_dict(d)
    it = make_sft_dataloader(args, toy_tokenize, 2, rank=0, world=1)
    b = next(it)
    assert set(b) == {"input_ids", "attention_mask", "labels"} and b["input_ids"].shape[0] == 4
    # prompt cut to 5 tokens, reply to 2 + eos: every row has exactly 8 real tokens, labels only on the last 3
    assert b["attention_mask"].sum(1).tolist() == [8] * 4 and ((b["labels"] != -100).sum(1) == 3).all()
    assert (b["labels"][:, -1] == 2).all()


def test_eight_simulated_ranks_cover_the_dataset_once():
    """mirrors the reference's tests/data/dataloader_test.py: 8 simulated ranks, every example is seen by exactly one rank
    (1000 examples, micro-batch 5: 1000 / 8 = 125 per rank = 25 full batches, nothing dropped)"""
    from dolomite_engine_b200.data.finetuning import batches

    n, world, mbs = 1000, 8, 5
    dataset = [{"input": [i, 1, 2], "output": [2]} for i in range(n)]
    seen = {}
    for rank in range(world):
        ids = []
        for batch in batches(dataset, mbs, eos_token_id=0, use_padding_free_transformer=True, rank=rank, world_size=world,
                             seed=42, infinite=False):
            ids += [row[0] for row in batch["input_ids"]]
        assert len(ids) == len(set(ids)), f"rank {rank} saw an example twice"
        seen[rank] = ids
    flat = [i for r in range(world) for i in seen[r]]
    assert len(flat) == n and set(flat) == set(range(n))
    # pretraining sampler: the same property per global batch
    from dolomite_engine_b200.data import MegatronBatchSampler

    rows = {r: [i for b in MegatronBatchSampler(960, 0, 4, world, r) for i in b] for r in range(world)}
    flat = sorted(i for r in rows for i in rows[r])
    assert flat == list(range(960)) and all(len(rows[r]) == 120 for r in rows)


def test_split_files_and_evaluate_loop(tmp_path):
    """validation split of the SFT feed (what `load_dataset(dir)[split]` resolves in the reference) and finetune.evaluate
    (finetune.py:156-219: mean loss over one pass, model back in training mode)"""
    import json
    import types

    import torch

    from dolomite_engine_b200.data.finetuning import JSONLinesSFTDataset, split_files
    from dolomite_engine_b200.finetune import evaluate, make_sft_val_batches

    d = tmp_path / "sft"
    d.mkdir()
    for name, n in (("train.jsonl", 6), ("validation.jsonl", 4), ("test.jsonl", 1)):
        with open(d / name, "w") as f:
            for i in range(n):
                f.write(json.dumps({"input": f"{name[0]}{i}", "output": "ok"}) + "\n")
    base = lambda p: [os.path.basename(x) for x in p]  # noqa: E731
    assert base(split_files(str(d), "train")) == ["train.jsonl"] and base(split_files(str(d), "val")) == ["validation.jsonl"]
    assert base(split_files(str(d), None)) == ["test.jsonl", "train.jsonl", "validation.jsonl"]
    plain = tmp_path / "plain"
    plain.mkdir()
    (plain / "a.jsonl").write_text(json.dumps({"input": "x", "output": "y"}) + "\n")
    assert base(split_files(str(plain), "train")) == ["a.jsonl"] and split_files(str(plain), "val") == []
    assert split_files(str(d / "train.jsonl"), "train") == [str(d / "train.jsonl")] and split_files(str(d / "train.jsonl"), "val") == []
    tok = lambda s: [ord(c) % 50 + 3 for c in s]  # noqa: E731
    assert len(JSONLinesSFTDataset(str(d), tok, 2, split="val")) == 4 and len(JSONLinesSFTDataset(str(d), tok, 2, split="train")) == 6

    ns = types.SimpleNamespace
    args = ns(datasets=[ns(class_args={"data_path": str(d)}, input_format="__input__", output_format="__output__",
                           max_input_tokens=None, max_output_tokens=None)],
              training_parameters=ns(eval_during_training=True, micro_batch_size=2, loss_mask="output_only"),
              model_args=ns(use_padding_free_transformer=True), random_args=ns(seed=1))
    val = make_sft_val_batches(args, tok, 2, rank=0, world=1)
    seen = []

    class Fake(torch.nn.Module):
        def forward(self, batch):
            seen.append((self.training, len(batch["input_ids"])))
            return torch.tensor(float(len(seen)))

    model = Fake()
    assert evaluate(val, model) == pytest.approx(1.5) and seen == [(False, 2), (False, 2)] and model.training
    assert evaluate(None, model) is None
    args.training_parameters.eval_during_training = False
    assert make_sft_val_batches(args, tok, 2, 0, 1) is None


def test_feed_resumes_where_it_stopped_and_ranks_draw_equal_counts():
    """a resumed finetuning run must continue with the next unseen batch (reference finetune.py:150, :286-305 hand
    the ResumableDataLoader to save / load), and with len(dataset) % world != 0 every rank still yields the same number of
    batches (rank-sharded evaluation issues one collective per batch)."""
    from dolomite_engine_b200.data.finetuning import batches
    from dolomite_engine_b200.pretrain import SyntheticPackedDataset

    ds = [{"input": [10 + i, 3], "output": [20 + i, 2]} for i in range(23)]  # 23 examples: not a multiple of 2 ranks x 2
    counts = []
    for rank in range(2):
        it = batches(ds, 2, 2, True, rank=rank, world_size=2, seed=5, infinite=False)
        counts.append(sum(1 for _ in it))
    assert counts[0] == counts[1] == 23 // 4

    a = batches(ds, 2, 2, True, rank=1, world_size=2, seed=5)
    seen = [next(a)["input_ids"] for _ in range(8)]  # crosses an epoch boundary (5 batches per epoch)
    state = a.state_dict()
    assert state["epoch"] == 1 and state["offset"] == 3
    expect = [next(a)["input_ids"] for _ in range(4)]
    b = batches(ds, 2, 2, True, rank=1, world_size=2, seed=5)
    b.load_state_dict(state)
    assert [next(b)["input_ids"] for _ in range(4)] == expect
    assert seen[0] != expect[0]

    s1 = SyntheticPackedDataset(100, 2, 8, rank=3, pin=False)
    first = [next(s1)["text"] for _ in range(5)]
    s2 = SyntheticPackedDataset(100, 2, 8, rank=3, pin=False)
    s2.load_state_dict({"consumed_samples": 3 * 2})
    assert torch.equal(next(s2)["text"], first[3]) and torch.equal(next(s2)["text"], first[4])
    assert s1.state_dict()["consumed_samples"] == 10
