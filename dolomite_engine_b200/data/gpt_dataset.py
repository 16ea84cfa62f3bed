"""Pretraining data feed: the producer of `{"text": int64[mbs, S+1]}` batches that sits immediately before the hot path
(SURVEY.md section 8f rank 2).  Same sample definition, shuffling and rank assignment as the reference
(data/megatron/gpt_dataset.py, blended_dataset.py, blended_megatron_dataset_builder.py, sampler.py, __init__.py), so a
run fed by this module sees exactly the token stream the reference would have seen with the same seed:

* `GPTDataset`      document index (epochs x documents, shuffled), sample index (one forward merge, native), shuffle index;
                    all three drawn from ONE `numpy.random.RandomState(seed)` in the reference's order (gpt_dataset.py:298-352)
* `BlendedDataset`  weighted mix of datasets through the greedy largest-deficit index (native)
* `MegatronBatchSampler`  rank r takes rows [r*mbs, (r+1)*mbs) of every global batch; resumable via `consumed_samples`
* `PackedBatchLoader`     assembles micro-batches with one native gather straight into PINNED host memory from the memory
                    mapped .bin (no per-sample numpy concatenation, no worker processes), one batch ahead of the trainer

* index cache     the three indices of a GPTDataset are stored / found under the reference's file names
                    (`<md5 of the unique description>-GPTDataset-{document,sample,shuffle}_index.npy`, gpt_dataset.py:265-330), so
                    caches written by either implementation are picked up by the other and every rank memory-maps ONE copy

Fill-in-the-middle (`fim_rate` > 0) rewrites rows on the host (data/fim.py).
"""

from __future__ import annotations

import ctypes
import hashlib
import json
import math
import os
import re
import threading
import warnings
from collections import OrderedDict
from queue import Queue

import numpy as np
import torch

from .indexed_dataset import MMapIndexedDataset

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    """the host helper library (built by dolomite_engine_b200.build); fails loudly when it is missing"""
    global _LIB
    if _LIB is None:
        path = os.path.join(os.path.dirname(_HERE), "lib", "libdolomite_data.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} not found: run `python -c 'import __graft_entry__ as g; g.build()'` first")
        lib = ctypes.CDLL(path)
        i64, p = ctypes.c_int64, ctypes.c_void_p
        lib.dolomite_data_build_sample_index_i32.restype = i64
        lib.dolomite_data_build_sample_index_i32.argtypes = [p, p, i64, i64, i64, i64, p]
        lib.dolomite_data_build_sample_index_i64.restype = i64
        lib.dolomite_data_build_sample_index_i64.argtypes = [p, p, i64, i64, i64, i64, p]
        lib.dolomite_data_num_samples.restype = i64
        lib.dolomite_data_num_samples.argtypes = [i64, i64, i64]
        lib.dolomite_data_build_blending_indices.restype = None
        lib.dolomite_data_build_blending_indices.argtypes = [p, p, p, ctypes.c_int32, i64]
        lib.dolomite_data_gather_rows.restype = ctypes.c_int32
        lib.dolomite_data_gather_rows.argtypes = [p, ctypes.c_int32, p, p, p, i64, i64, p]
        _LIB = lib
    return _LIB


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


# ------------------------------------------------------------------------------------------------
# index builders
# ------------------------------------------------------------------------------------------------
def build_sample_index(sizes: np.ndarray, document_index: np.ndarray, sequence_length: int, num_epochs: int,
                       tokens_per_epoch: int) -> np.ndarray:
    """[num_samples + 1, 2] (slot in document_index, token offset) -- helpers.cpp build_sample_idx_int32/int64; dtype follows
    the document index like the reference's dispatch (utils/__init__.py)"""
    sizes = np.ascontiguousarray(sizes, dtype=np.int32)
    lib = _lib()
    rows = int(lib.dolomite_data_num_samples(sequence_length, num_epochs, int(tokens_per_epoch))) + 1
    if document_index.dtype == np.int32:
        di = np.ascontiguousarray(document_index)
        out = np.empty((rows, 2), dtype=np.int32)
        n = lib.dolomite_data_build_sample_index_i32(_ptr(sizes), _ptr(di), di.size, sequence_length, num_epochs,
                                                     int(tokens_per_epoch), _ptr(out))
    else:
        di = np.ascontiguousarray(document_index, dtype=np.int64)
        out = np.empty((rows, 2), dtype=np.int64)
        n = lib.dolomite_data_build_sample_index_i64(_ptr(sizes), _ptr(di), di.size, sequence_length, num_epochs,
                                                     int(tokens_per_epoch), _ptr(out))
    if n != rows:
        raise ValueError("document index holds fewer tokens than num_epochs * tokens_per_epoch")
    return out


def build_blending_indices(weights: list[float], size: int) -> tuple[np.ndarray, np.ndarray]:
    """(dataset_index int16[size], dataset_sample_index int64[size]) -- helpers.cpp build_blending_indices"""
    w = np.ascontiguousarray(weights, dtype=np.float64)
    di = np.zeros(size, dtype=np.int16)
    dsi = np.zeros(size, dtype=np.int64)
    _lib().dolomite_data_build_blending_indices(_ptr(di), _ptr(dsi), _ptr(w), len(weights), size)
    return di, dsi


def get_num_epochs(num_tokens_per_epoch: int, seq_length: int, num_samples: int) -> int:
    """gpt_dataset.py:416-435: smallest e with (e * tokens_per_epoch - 1) // S >= num_samples"""
    need = num_samples * seq_length + 1
    return max(1, -(-need // int(num_tokens_per_epoch)))


def build_document_index(documents: np.ndarray, num_epochs: int, rs: np.random.RandomState, separate_final_epoch: bool):
    """gpt_dataset.py:438-468"""
    if not separate_final_epoch or num_epochs == 1:
        idx = np.tile(documents, num_epochs).astype(documents.dtype)
        rs.shuffle(idx)
        return idx
    first = build_document_index(documents, num_epochs - 1, rs, False)
    last = build_document_index(documents, 1, rs, False)
    return np.concatenate((first, last))


def build_shuffle_index(num_samples: int, total_size: int, rs: np.random.RandomState) -> np.ndarray:
    """gpt_dataset.py:471-499"""
    dtype = np.uint32 if total_size < np.iinfo(np.uint32).max - 1 else np.int64
    first = np.arange(0, num_samples, dtype=dtype)
    rs.shuffle(first)
    if num_samples == total_size:
        return first
    last = np.arange(num_samples, total_size, dtype=dtype)
    rs.shuffle(last)
    return np.concatenate((first, last))


def parse_and_normalize_split(split: str) -> list[float]:
    """blended_megatron_dataset_config.py:98-114"""
    v = [float(x) for x in re.findall(r"[.0-9]+", split)]
    v = v + [0.0] * (3 - len(v))
    assert len(v) == 3 and all(x >= 0 for x in v)
    s = sum(v)
    return [x / s for x in v]


def get_split_indices(split: list[float], num_elements: int) -> list[int]:
    """blended_megatron_dataset_builder.py:376-397"""
    idx = [0]
    for pct in split:
        idx.append(idx[-1] + int(round(pct * float(num_elements))))
    over = idx[-1] - num_elements
    idx[1:] = [x - over for x in idx[1:]]
    assert idx[-1] == num_elements
    return idx


def get_train_val_test_samples(num_training_steps: int, micro_batch_size: int, gradient_accumulation_steps: int,
                               eval_interval: int | None, eval_steps: int | None, dp_world_size: int) -> tuple[int, int, int]:
    """data/megatron/__init__.py:215-234"""
    per_step = micro_batch_size * gradient_accumulation_steps * dp_world_size
    train = num_training_steps * per_step
    if not eval_interval or not eval_steps:
        return train, 0, 0
    return train, (num_training_steps // eval_interval + 1) * eval_steps * per_step, eval_steps * per_step


# ------------------------------------------------------------------------------------------------
# datasets
# ------------------------------------------------------------------------------------------------
class GPTDataset:
    """Samples of S+1 tokens cut from the shuffled, epoch-repeated document stream (gpt_dataset.py:30-400)"""

    def __init__(self, indexed_dataset: MMapIndexedDataset, indexed_indices: np.ndarray, num_samples: int,
                 sequence_length: int, random_seed: int = 1234, fim=None, *, index_split: str = "train",
                 split: str | None = None, name: str | None = None, path_to_cache: str | None = None, cache: str = "off"):
        # fim: data.fim.FIMSpec or None; its random stream is seeded like the reference's (gpt_dataset.py:55)
        self.fim = fim if (fim is not None and fim.rate != 0) else None
        self.np_rng = np.random.RandomState(seed=random_seed)
        self.indexed_dataset = indexed_dataset
        self.indexed_indices = np.asarray(indexed_indices)
        self.num_samples = int(num_samples)
        self.sequence_length = int(sequence_length)
        self.random_seed = random_seed
        # what identifies the indices on disk: megatron_dataset.py:52-61 (`_key_config_attributes` = name, split, random_seed,
        # sequence_length of the GPTDatasetConfig) -- same keys, same order, same JSON layout, hence the same MD5 and file names
        ident = OrderedDict()
        ident["class"] = "GPTDataset"
        ident["path_prefix"] = indexed_dataset.path_prefix
        ident["num_samples"] = self.num_samples
        ident["index_split"] = index_split  # Split.{train,valid,test}.name
        ident["name"], ident["split"] = name, split
        ident["random_seed"], ident["sequence_length"] = random_seed, self.sequence_length
        self.unique_description = json.dumps(ident, indent=4)
        self.unique_description_hash = hashlib.md5(self.unique_description.encode("utf-8")).hexdigest()
        if cache not in ("off", "load", "build"):
            raise ValueError(f"cache={cache!r}: off (build in memory), load (use stored indices when present, never write) or "
                             "build (store them when missing, like the reference's caching_allowed ranks)")
        self.path_to_cache, self.cache_mode, self.cache_hit = path_to_cache, cache, False
        self.document_index, self.sample_index, self.shuffle_index = self._cached_indices()

    def cache_paths(self) -> dict[str, str]:
        """gpt_dataset.py:265-275: default directory `<path_prefix>/cache/GPTDataset_indices`"""
        root = self.path_to_cache
        if root is None:
            root = os.path.join(self.indexed_dataset.path_prefix, "cache", "GPTDataset_indices")
        return {k: os.path.join(root, f"{self.unique_description_hash}-GPTDataset-{k}")
                for k in ("description.txt", "document_index.npy", "sample_index.npy", "shuffle_index.npy")}

    def _cached_indices(self):
        """gpt_dataset.py:241-400.  `build`: a miss builds the indices, stores them (description first, every array through a
        temporary file + rename so that another node never maps half a file) and maps them back read-only; `load`: stored
        indices are used when all four files exist, nothing is ever written; `off`: always in memory.  A cache directory that
        cannot be written degrades to the in-memory indices with a warning (the reference raises there)."""
        if self.cache_mode == "off":
            return self._build_indices()
        paths = self.cache_paths()
        hit = all(os.path.isfile(v) for v in paths.values())
        built = None
        if not hit:
            built = self._build_indices()
            if self.cache_mode == "load":
                return built
            try:
                os.makedirs(os.path.dirname(paths["description.txt"]), exist_ok=True)
                with open(paths["description.txt"], "wt") as f:
                    f.write(self.unique_description)
                for key, arr in zip(("document_index.npy", "sample_index.npy", "shuffle_index.npy"), built):
                    tmp = f"{paths[key]}.tmp{os.getpid()}.npy"
                    np.save(tmp, arr, allow_pickle=True)
                    os.replace(tmp, paths[key])
            except OSError as e:
                warnings.warn(f"GPTDataset indices not cached under {os.path.dirname(paths['description.txt'])} ({e}); "
                              "set class_args.data_cache_path to a writable directory")
                return built
        else:
            self.cache_hit = True
            self._set_epoch_counters()
        loaded = tuple(np.load(paths[k], allow_pickle=True, mmap_mode="r")
                       for k in ("document_index.npy", "sample_index.npy", "shuffle_index.npy"))
        if built is not None:  # what was just written must read back as what was built
            assert all(a.shape == b.shape and a.dtype == b.dtype for a, b in zip(built, loaded))
        return loaded

    def _set_epoch_counters(self) -> None:
        sizes = self.indexed_dataset.sequence_lengths
        self.tokens_per_epoch = int(np.sum(sizes[self.indexed_indices]))
        self.num_epochs = get_num_epochs(self.tokens_per_epoch, self.sequence_length, self.num_samples)

    def _build_indices(self):
        sizes = self.indexed_dataset.sequence_lengths
        tokens_per_epoch = int(np.sum(sizes[self.indexed_indices]))
        S = self.sequence_length
        num_epochs = get_num_epochs(tokens_per_epoch, S, self.num_samples)
        if num_epochs == 1:
            separate_final_epoch = False
            sans_final = None
        else:
            sans_final = ((num_epochs - 1) * tokens_per_epoch - 1) // S
            from_final = self.num_samples - sans_final
            per_epoch = (tokens_per_epoch - 1) // S
            assert 0 <= from_final <= per_epoch + 1
            separate_final_epoch = from_final < int(0.80 * per_epoch)
        rs = np.random.RandomState(self.random_seed)
        document_index = build_document_index(self.indexed_indices, num_epochs, rs, separate_final_epoch)
        sample_index = build_sample_index(sizes, document_index, S, num_epochs, tokens_per_epoch)
        total = sample_index.shape[0] - 1
        shuffle_index = build_shuffle_index(sans_final if separate_final_epoch else total, total, rs)
        self.num_epochs, self.tokens_per_epoch = num_epochs, tokens_per_epoch
        return document_index, sample_index, shuffle_index

    def __len__(self) -> int:
        return self.sample_index.shape[0] - 1

    def sample_parts(self, idx: int) -> list[tuple[int, int]]:
        """(element offset into .bin, length) of every document slice of sample idx (gpt_dataset.py:117-160)"""
        i = int(self.shuffle_index[idx])
        d0, o0 = (int(x) for x in self.sample_index[i])
        d1, o1 = (int(x) for x in self.sample_index[i + 1])
        offs = self.indexed_dataset.sequence_element_offsets
        sizes = self.indexed_dataset.sequence_lengths
        if d0 == d1:
            doc = int(self.document_index[d0])
            return [(int(offs[doc]) + o0, o1 - o0 + 1)]
        parts = []
        for d in range(d0, d1 + 1):
            doc = int(self.document_index[d])
            off = o0 if d == d0 else 0
            length = (o1 + 1 if d == d1 else int(sizes[doc])) - off
            parts.append((int(offs[doc]) + off, length))
        return parts

    def __getitem__(self, idx: int) -> dict[str, np.ndarray]:
        tok = self.indexed_dataset.tokens
        text = np.concatenate([tok[o : o + n] for o, n in self.sample_parts(idx)]).astype(np.int64)
        if self.fim is not None:
            from .fim import apply_fim

            text = apply_fim(text, self.np_rng, self.fim)
        return {"text": text}


class BlendedDataset:
    """Weighted mix (blended_dataset.py:26-166): sample i comes from datasets[dataset_index[i]][dataset_sample_index[i]]"""

    def __init__(self, datasets: list[GPTDataset], weights: list[float], size: int):
        assert len(datasets) == len(weights) and len(datasets) < np.iinfo(np.int16).max
        s = float(sum(weights))
        self.weights = [w / s for w in weights]
        self.datasets, self.size = datasets, int(size)
        self.dataset_index, self.dataset_sample_index = build_blending_indices(self.weights, self.size)
        for d, ds in enumerate(datasets):
            need = int((self.dataset_index == d).sum())
            if need > len(ds):
                raise IndexError(f"blend needs {need} samples of dataset {d} which holds {len(ds)}")

    def __len__(self) -> int:
        return self.size

    def locate(self, idx: int) -> tuple[GPTDataset, int]:
        if not 0 <= idx < self.size:
            raise IndexError(idx)
        return self.datasets[int(self.dataset_index[idx])], int(self.dataset_sample_index[idx])

    def __getitem__(self, idx: int) -> dict:
        ds, j = self.locate(idx)
        return {"dataset_id": int(self.dataset_index[idx]), **ds[j]}


def build_gpt_datasets(data_path, split: str | None, sizes: tuple[int, int, int], sequence_length: int, seed: int, fim=None, *,
                       blend_per_split: list | None = None, data_cache_path: str | None = None, cache: str = "off",
                       node_uses_local_storage: bool = False):
    """Options 1-3 of data/megatron/__init__.py:93-101 -> (train, val, test), each a GPTDataset / BlendedDataset / None
    (blended_megatron_dataset_builder.py:61-226):
      1 / 2  `data_path` = one prefix, or [w1, prefix1, w2, prefix2, ...], cut into the three splits by `split` ("98,1,1");
      3      `blend_per_split` = [train blend, validation blend, test blend] (class_args train_data_path / val_data_path /
             test_data_path): every split has its own stores and uses them whole; `data_path` must be None and `split` is ignored.

    `cache` = "build" follows blended_megatron_dataset_builder.py:330-366 under torch.distributed: rank 0 (and local rank 0 of every
    node with `node_uses_local_storage`) builds and stores the indices, everybody meets at a barrier, the other ranks then find
    them (mode "load": a rank that still misses them builds in memory instead of failing)."""
    args = (data_path, split, sizes, sequence_length, seed, fim, data_cache_path)
    if cache == "build":
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            allowed = dist.get_rank() == 0 or (node_uses_local_storage and int(os.environ.get("LOCAL_RANK", "0")) == 0)
            out = _build_gpt_datasets(*args, "build", blend_per_split) if allowed else None
            dist.barrier()
            if not allowed:
                out = _build_gpt_datasets(*args, "load", blend_per_split)
            return out
    return _build_gpt_datasets(*args, cache, blend_per_split)


_SPLIT_NAMES = ("train", "valid", "test")  # data/megatron/utils/__init__.py:14-17 `Split`


def _build_gpt_datasets(data_path, split, sizes, sequence_length: int, seed: int, fim, data_cache_path, cache: str,
                        blend_per_split=None):
    def one(prefix: str, split_v: list[float], want: list[int], split_desc):
        """_build_megatron_dataset_splits: the three GPTDatasets (or None) cut from ONE store"""
        ids = MMapIndexedDataset(prefix)
        bounds = get_split_indices(split_v, ids.sequence_lengths.shape[0])
        dt = np.int32 if max(bounds) <= np.iinfo(np.int32).max else np.int64
        out = []
        for i in range(3):
            if split_v[i] == 0.0 or want[i] == 0:
                out.append(None)
            else:
                out.append(GPTDataset(ids, np.arange(bounds[i], bounds[i + 1], dtype=dt), want[i], sequence_length, seed, fim=fim,
                                      index_split=_SPLIT_NAMES[i], split=split_desc, path_to_cache=data_cache_path, cache=cache))
        return out

    def blended(blend: list, split_v: list[float], want: list[int], split_desc) -> list:
        """one blend (a single prefix or weight / prefix pairs) -> the three splits"""
        if len(blend) == 1:
            return one(blend[0], split_v, list(want), split_desc)
        assert len(blend) % 2 == 0, "a blend is one prefix or [weight, prefix, weight, prefix, ...]"
        weights = [float(blend[i]) for i in range(0, len(blend), 2)]
        prefixes = [str(blend[i]).strip() for i in range(1, len(blend), 2)]
        tot = sum(weights)
        weights = [w / tot for w in weights]
        per = [[int(math.ceil(n * w * 1.005)) for n in want] for w in weights]  # 0.5 % margin like the reference
        parts = [one(p, split_v, per[k], split_desc) for k, p in enumerate(prefixes)]
        res = []
        for i in range(3):
            dss = [parts[k][i] for k in range(len(prefixes))]
            # the blend's length is the SUM of the per-store requests (margin included), blended_megatron_dataset_builder.py:98-117
            res.append(None if any(d is None for d in dss) else BlendedDataset(dss, weights, sum(per[k][i] for k in range(len(prefixes)))))
        return res

    if blend_per_split is not None and any(blend_per_split):
        # blended_megatron_dataset_config.py:68-74 / builder :122-170: the split string is dropped (the description stores null)
        assert data_path is None, "blend (data_path) and blend_per_split (train / val / test data paths) are incompatible"
        assert len(blend_per_split) == 3, "blend_per_split must contain 3 blends"
        out = []
        for i in range(3):
            blend = blend_per_split[i]
            if not blend:
                out.append(None)
                continue
            blend = [blend] if isinstance(blend, str) else list(blend)
            split_spoof, sizes_spoof = [0.0] * 3, [0] * 3
            split_spoof[i], sizes_spoof[i] = 1.0, sizes[i]
            out.append(blended(blend, split_spoof, sizes_spoof, None)[i])
        return tuple(out)
    assert split is not None, "both blend and split must be provided"
    if isinstance(data_path, str):
        data_path = [data_path]
    return tuple(blended(list(data_path), parse_and_normalize_split(split), list(sizes), split))


# ------------------------------------------------------------------------------------------------
# sampler + loader
# ------------------------------------------------------------------------------------------------
class MegatronBatchSampler:
    """sampler.py:4-47: consecutive global batches of mbs * num_replicas samples; rank r takes rows [r*mbs, (r+1)*mbs)"""

    def __init__(self, total_samples: int, consumed_samples: int, micro_batch_size: int, num_replicas: int, rank: int,
                 drop_last: bool = True):
        assert total_samples > 0, f"no sample to consume: {total_samples}"
        assert consumed_samples < total_samples, f"no samples left to consume: {consumed_samples}, {total_samples}"
        assert micro_batch_size > 0
        self.total_samples, self.consumed_samples = total_samples, consumed_samples
        self.micro_batch_size, self.num_replicas, self.rank, self.drop_last = micro_batch_size, num_replicas, rank, drop_last

    def __len__(self) -> int:
        return self.total_samples

    def __iter__(self):
        g = self.micro_batch_size * self.num_replicas
        lo = self.rank * self.micro_batch_size
        start = self.consumed_samples
        while start + g <= self.total_samples:
            yield list(range(start + lo, start + lo + self.micro_batch_size))
            start += g
        if start < self.total_samples and not self.drop_last:
            tail = list(range(start, self.total_samples))
            yield tail[lo : lo + self.micro_batch_size]


class PackedBatchLoader:
    """Iterator of `{"text": LongTensor[mbs, S+1]}` (pinned when CUDA is present): what `ModelWrapperForPretraining.forward`
    consumes (model_wrapper/pretraining.py:89).  A background thread keeps `prefetch` batches assembled ahead of the trainer;
    every batch is one native gather from the memory-mapped token files into the pinned buffer."""

    def __init__(self, dataset, sampler: MegatronBatchSampler, sequence_length: int, prefetch: int = 2, pin: bool | None = None):
        self.dataset, self.sampler, self.row_len = dataset, sampler, sequence_length + 1
        self.pin = torch.cuda.is_available() if pin is None else pin
        self.prefetch = max(1, prefetch)
        self.consumed_samples = sampler.consumed_samples

    def _assemble(self, rows: list[int]) -> torch.Tensor:
        out = torch.empty((len(rows), self.row_len), dtype=torch.int64, pin_memory=self.pin)
        dst = out.numpy()
        # group rows by backing token file (a blend mixes several), one native call per file
        by_file: dict[int, tuple[MMapIndexedDataset, list[int], list[list[tuple[int, int]]]]] = {}
        for r, idx in enumerate(rows):
            ds, j = self.dataset.locate(idx) if hasattr(self.dataset, "locate") else (self.dataset, idx)
            key = id(ds.indexed_dataset)
            by_file.setdefault(key, (ds.indexed_dataset, [], []))
            by_file[key][1].append(r)
            by_file[key][2].append(ds.sample_parts(j))
        lib = _lib()
        for ids, rws, parts in by_file.values():
            flat = [p for ps in parts for p in ps]
            part_ptr = np.asarray([p[0] for p in flat], dtype=np.int64)
            part_len = np.asarray([p[1] for p in flat], dtype=np.int64)
            first = np.zeros(len(rws) + 1, dtype=np.int64)
            np.cumsum([len(ps) for ps in parts], out=first[1:])
            tmp = dst if len(by_file) == 1 else np.empty((len(rws), self.row_len), dtype=np.int64)
            rc = lib.dolomite_data_gather_rows(ids.tokens.ctypes.data, ids.dtype.itemsize, _ptr(part_ptr), _ptr(part_len),
                                               _ptr(first), len(rws), self.row_len, _ptr(tmp))
            if rc != 0:
                raise RuntimeError(f"gather_rows failed ({rc}): sample parts do not add up to {self.row_len} tokens")
            if tmp is not dst:
                dst[rws] = tmp
        # fill-in-the-middle rewrites rows on the host, in row order (each dataset owns its random stream)
        for r, idx in enumerate(rows):
            ds, _ = self.dataset.locate(idx) if hasattr(self.dataset, "locate") else (self.dataset, idx)
            if getattr(ds, "fim", None) is not None:
                from .fim import apply_fim

                dst[r] = apply_fim(dst[r].copy(), ds.np_rng, ds.fim)
        return out

    def __iter__(self):
        q: Queue = Queue(maxsize=self.prefetch)
        stop = threading.Event()

        def work():
            try:
                for rows in self.sampler:
                    if stop.is_set():
                        return
                    q.put(self._assemble(rows))
                q.put(None)
            except BaseException as e:  # surface producer errors in the consumer
                q.put(e)

        t = threading.Thread(target=work, daemon=True)
        t.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                self.consumed_samples += self.sampler.micro_batch_size * self.sampler.num_replicas
                yield {"text": item}
        finally:
            stop.set()
