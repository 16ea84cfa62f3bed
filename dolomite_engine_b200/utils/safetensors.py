"""Safetensors reader/writer with the interface of the reference's SafeTensorsWeightsManager
(utils/safetensors.py:11-97): lazy per-tensor access over one or many *.safetensors files, sharded save."""

from __future__ import annotations

import json
import os

import torch
from safetensors import safe_open
from safetensors.torch import save_file

SAFE_WEIGHTS_INDEX_NAME = "model.safetensors.index.json"
_MAX_SHARD_BYTES = 5 * 1024**3


class SafeTensorsWeightsManager:
    def __init__(self, model_path: str) -> None:
        if model_path.endswith(".safetensors"):
            filenames = [model_path]
        else:
            filenames = sorted(os.path.join(model_path, f) for f in os.listdir(model_path) if f.endswith(".safetensors"))
        self.tensor_filenames: dict[str, str] = {}
        self.file_handles = {}
        for filename in filenames:
            f = safe_open(filename, framework="pytorch")
            self.file_handles[filename] = f
            for tensor_name in f.keys():
                self.tensor_filenames[tensor_name] = filename

    def get_slice(self, tensor_name: str):
        return self.file_handles[self.tensor_filenames[tensor_name]].get_slice(tensor_name)

    def get_tensor(self, tensor_name: str, dtype: torch.dtype | None = None, device=None) -> torch.Tensor:
        t = self.file_handles[self.tensor_filenames[tensor_name]].get_tensor(tensor_name)
        return t.to(dtype=dtype, device=device)

    def get_shape(self, tensor_name: str):
        return self.get_slice(tensor_name).get_shape()

    def has_tensor(self, tensor_name: str) -> bool:
        return tensor_name in self.tensor_filenames

    def __len__(self) -> int:
        return len(self.tensor_filenames)

    def __iter__(self):
        yield from self.tensor_filenames

    def __eq__(self, other: object) -> bool:
        if not isinstance(other, SafeTensorsWeightsManager) or len(self) != len(other):
            return False
        for a, b in zip(self, other):
            if a != b or not self.get_tensor(a).equal(other.get_tensor(b)):
                return False
        return True

    def state_dict(self) -> dict:
        return {name: self.get_tensor(name) for name in self}

    @staticmethod
    def save_state_dict(state_dict: dict, save_path: str) -> None:
        os.makedirs(save_path, exist_ok=True)
        shards: list[dict] = [{}]
        size = 0
        for name, t in state_dict.items():
            nbytes = t.numel() * t.element_size()
            if size + nbytes > _MAX_SHARD_BYTES and shards[-1]:
                shards.append({})
                size = 0
            shards[-1][name] = t.contiguous()
            size += nbytes
        if len(shards) == 1:
            save_file(shards[0], os.path.join(save_path, "model.safetensors"), metadata={"format": "pt"})
            return
        weight_map = {}
        total = 0
        for i, shard in enumerate(shards):
            fname = f"model-{i + 1:05d}-of-{len(shards):05d}.safetensors"
            save_file(shard, os.path.join(save_path, fname), metadata={"format": "pt"})
            for name, t in shard.items():
                weight_map[name] = fname
                total += t.numel() * t.element_size()
        with open(os.path.join(save_path, SAFE_WEIGHTS_INDEX_NAME), "w") as f:
            f.write(json.dumps({"metadata": {"total_size": total}, "weight_map": weight_map}, indent=2))
