"""Supervised finetuning feed for decoder-only models (the producer of `ModelWrapperForFinetuning`'s batches; reference:
data/base.py:84-121 example construction, data/utils.py:8-92 `collate_fn`, data/instruction_tuning/base.py JSONL reader).

An example is `{"input": prompt + response + eos, "output": response + eos}` (token ids); a batch is either
  * padding-free (`use_padding_free_transformer: true`):  {"input_ids": list[list[int]], "labels": list[list[int]]}
  * padded, LEFT padding with eos:                          {"input_ids", "attention_mask", "labels"} LongTensors [B, S]
with `labels = -100` on the prompt (loss_mask output_only) or equal to the inputs (no_mask).  Encoder-decoder models are
not part of the B200 path.
"""

from __future__ import annotations

import json
import os
from typing import Callable, Iterable

import torch

LABELS_MASK_VALUE = -100


def build_example(tokenize: Callable[[str], list[int]], eos_token_id: int, input_text: str, output_text: str | None,
                  max_input_tokens: int | None = None, max_output_tokens: int | None = None) -> dict:
    """data/base.py:84-121 (decoder-only): prompt truncated to max_input_tokens, response to max_output_tokens - 1, then eos"""
    prompt = list(tokenize(input_text))
    if max_input_tokens is not None:
        prompt = prompt[:max_input_tokens]
    if output_text is None:  # inference mode
        return {"input": prompt}
    response = list(tokenize(output_text))
    if max_output_tokens is not None:
        response = response[: max_output_tokens - 1]
    response.append(eos_token_id)
    return {"input": prompt + response, "output": response}


def collate(batch: list[dict], eos_token_id: int, use_padding_free_transformer: bool, loss_mask: str = "output_only",
            training: bool = True) -> dict:
    """data/utils.py:8-92, decoder-only branch"""
    if loss_mask not in ("output_only", "no_mask"):
        raise ValueError(f"unexpected loss_mask ({loss_mask})")
    inputs = [ex["input"] for ex in batch]
    outputs = [ex["output"] for ex in batch] if training else None
    if use_padding_free_transformer:
        out = {"input_ids": inputs}
        if training:
            out["labels"] = inputs if loss_mask == "no_mask" else [
                [LABELS_MASK_VALUE] * (len(i) - len(o)) + list(o) for i, o in zip(inputs, outputs)]
        return out
    width = max(len(i) for i in inputs)
    pad = lambda row, fill: [fill] * (width - len(row)) + list(row)  # noqa: E731  (left padding)
    out = {"input_ids": torch.tensor([pad(i, eos_token_id) for i in inputs], dtype=torch.long),
           "attention_mask": torch.tensor([pad([1] * len(i), 0) for i in inputs], dtype=torch.long)}
    if training:
        rows = inputs if loss_mask == "no_mask" else outputs
        out["labels"] = torch.tensor([pad(r, LABELS_MASK_VALUE) for r in rows], dtype=torch.long)
    return out


_SPLIT_PREFIXES = {"train": ("train",), "val": ("val", "validation", "dev"), "test": ("test",)}


def split_files(data_path: str, split: str | None) -> list[str]:
    """JSON-lines files of one split.  A directory whose files are named after splits (`train*.jsonl`, `val*` /
    `validation*` / `dev*`, `test*` -- what `datasets.load_dataset(dir)[split]` resolves in the reference,
    data/huggingface.py:46-47) is divided accordingly; any other directory, or a single file, is all training data."""
    if os.path.isfile(data_path):
        return [data_path] if split in (None, "train") else []
    names = sorted(f for f in os.listdir(data_path) if f.endswith((".jsonl", ".json")))
    kind = {f: next((s for s, pre in _SPLIT_PREFIXES.items() if f.lower().startswith(pre)), None) for f in names}
    if split is None or not any(kind.values()):
        return [os.path.join(data_path, f) for f in names] if split in (None, "train") else []
    return [os.path.join(data_path, f) for f in names if kind[f] == split]


class JSONLinesSFTDataset:
    """data/instruction_tuning/base.py: every line of every `*.jsonl` under `data_path` is {"input": str, "output": str};
    `input_format` / `output_format` wrap the raw strings ("__input__" / "__output__" placeholders, data/base.py:56-82)"""

    def __init__(self, data_path: str, tokenize: Callable[[str], list[int]], eos_token_id: int,
                 input_format: str = "__input__", output_format: str = "__output__", max_input_tokens: int | None = None,
                 max_output_tokens: int | None = None, training: bool = True, split: str | None = None):
        files = split_files(data_path, split)
        self.examples = []
        for path in files:
            with open(path) as fh:
                for line in fh:
                    if not line.strip():
                        continue
                    raw = json.loads(line)
                    text_in = input_format.replace("__input__", raw["input"])
                    text_out = output_format.replace("__output__", raw["output"]) if training else None
                    self.examples.append(build_example(tokenize, eos_token_id, text_in, text_out, max_input_tokens,
                                                       max_output_tokens))

    def __len__(self) -> int:
        return len(self.examples)

    def __getitem__(self, i: int) -> dict:
        return self.examples[i]


class ResumableBatches:
    """shuffled, rank-sharded iterator of collated micro-batches (BlendedDistributedSampler + ResumableDataLoader of the
    reference, data/dataloader.py / data/sampler.py, reduced to what train_step needs): epoch e uses permutation
    `randperm(len, seed + e)` truncated to a multiple of `world_size * micro_batch_size` (every rank yields the SAME number
    of batches, so rank-sharded evaluation issues the same collectives everywhere), rank r takes every world_size-th example
    starting at r.  `state_dict()` / `load_state_dict()` carry (epoch, batches consumed in the epoch): a resumed run continues
    with the batch the interrupted run would have drawn next (finetune.py:150 / :286-305 pass the loader to the checkpoint)."""

    def __init__(self, dataset, micro_batch_size: int, eos_token_id: int, use_padding_free_transformer: bool, rank: int = 0,
                 world_size: int = 1, seed: int = 42, loss_mask: str = "output_only", infinite: bool = True):
        self.dataset, self.mbs, self.eos = dataset, micro_batch_size, eos_token_id
        self.padding_free, self.rank, self.world = use_padding_free_transformer, rank, world_size
        self.seed, self.loss_mask, self.infinite = seed, loss_mask, infinite
        self.epoch, self.offset = 0, 0  # offset: micro-batches of this rank already drawn in `epoch`
        self._order: list[int] | None = None

    def batches_per_epoch(self) -> int:
        return len(self.dataset) // (self.world * self.mbs)

    def _epoch_order(self) -> list[int]:
        order = torch.randperm(len(self.dataset), generator=torch.Generator().manual_seed(self.seed + self.epoch)).tolist()
        usable = self.batches_per_epoch() * self.world * self.mbs
        return order[:usable][self.rank :: self.world]

    def __iter__(self):
        return self

    def __next__(self) -> dict:
        n = self.batches_per_epoch()
        if n == 0:
            raise StopIteration
        if self.offset >= n:
            if not self.infinite:
                raise StopIteration
            self.epoch, self.offset, self._order = self.epoch + 1, 0, None
        if self._order is None:
            self._order = self._epoch_order()
        lo = self.offset * self.mbs
        self.offset += 1
        return collate([self.dataset[j] for j in self._order[lo : lo + self.mbs]], self.eos, self.padding_free, self.loss_mask)

    def state_dict(self) -> dict:
        return {"epoch": self.epoch, "offset": self.offset, "seed": self.seed, "world_size": self.world,
                "micro_batch_size": self.mbs}

    def load_state_dict(self, state: dict) -> None:
        if state.get("world_size", self.world) != self.world or state.get("micro_batch_size", self.mbs) != self.mbs:
            raise ValueError("the finetuning feed can only be resumed with the world size / micro batch size it was saved with")
        self.epoch, self.offset, self._order = int(state["epoch"]), int(state["offset"]), None


def batches(dataset, micro_batch_size: int, eos_token_id: int, use_padding_free_transformer: bool, rank: int = 0,
            world_size: int = 1, seed: int = 42, loss_mask: str = "output_only", infinite: bool = True) -> ResumableBatches:
    return ResumableBatches(dataset, micro_batch_size, eos_token_id, use_padding_free_transformer, rank, world_size, seed,
                            loss_mask, infinite)
