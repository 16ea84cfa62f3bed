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
        paths = [model_path] if model_path.endswith(".safetensors") else sorted(
            os.path.join(model_path, f) for f in os.listdir(model_path) if f.endswith(".safetensors"))
        self._files = {p: safe_open(p, framework="pytorch") for p in paths}
        self._where: dict[str, str] = {name: p for p, handle in self._files.items() for name in handle.keys()}

    def _handle(self, tensor_name: str):
        return self._files[self._where[tensor_name]]

    def get_slice(self, tensor_name: str):
        return self._handle(tensor_name).get_slice(tensor_name)

    def get_tensor(self, tensor_name: str, dtype: torch.dtype | None = None, device=None) -> torch.Tensor:
        return self._handle(tensor_name).get_tensor(tensor_name).to(dtype=dtype, device=device)

    def get_shape(self, tensor_name: str):
        return self.get_slice(tensor_name).get_shape()

    def has_tensor(self, tensor_name: str) -> bool:
        return tensor_name in self._where

    def __len__(self) -> int:
        return len(self._where)

    def __iter__(self):
        return iter(self._where)

    def __eq__(self, other: object) -> bool:
        """same tensor names in the same order, same values"""
        if not isinstance(other, SafeTensorsWeightsManager) or list(self) != list(other):
            return False
        return all(self.get_tensor(name).equal(other.get_tensor(name)) for name in self)

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
