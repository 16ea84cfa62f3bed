"""Megatron memory-mapped token store: `<prefix>.bin` (raw token ids) + `<prefix>.idx` (lengths, byte pointers, document
boundaries).  Byte-for-byte the on-disk format of the reference (data/megatron/indexed_dataset.py:26-224 writer,
:226-338 reader, :340-523 dataset, :525-609 builder), so corpora tokenised for dolomite-engine / Megatron-LM are read
as they are and files written here are readable by the reference.

.idx layout (little endian):
    9 B   magic  b"MMIDIDX\\x00\\x00"
    u64   version = 1
    u8    dtype code (1 uint8, 2 int8, 3 int16, 4 int32, 5 int64, 6 float64, 7 float32, 8 uint16)
    u64   sequence count N
    u64   document count D  (entries of the document index, normally n_documents + 1)
    i32   sequence_lengths[N]
    i64   sequence_pointers[N]   byte offset of every sequence inside .bin
    i64   document_indices[D]    sequence index at which every document starts (last = N)
    i8    sequence_modes[N]      only for multimodal stores
"""

from __future__ import annotations

import os
import shutil
import struct

import numpy as np

_INDEX_HEADER = b"MMIDIDX\x00\x00"
_DTYPE_BY_CODE = {1: np.uint8, 2: np.int8, 3: np.int16, 4: np.int32, 5: np.int64, 6: np.float64, 7: np.float32, 8: np.uint16}
_CODE_BY_DTYPE = {np.dtype(v): k for k, v in _DTYPE_BY_CODE.items()}


def get_idx_path(path_prefix: str) -> str:
    return path_prefix + ".idx"


def get_bin_path(path_prefix: str) -> str:
    return path_prefix + ".bin"


def optimal_dtype(cardinality: int | None):
    """indexed_dataset.py:86-98: uint16 token ids when the vocabulary fits, else int32"""
    return np.uint16 if cardinality is not None and cardinality < 65500 else np.int32


class _Index:
    """parsed .idx (all arrays are views of one read-only memory map)"""

    def __init__(self, idx_path: str, multimodal: bool = False):
        with open(idx_path, "rb") as f:
            magic = f.read(9)
            if magic != _INDEX_HEADER:
                raise ValueError(f"bad header in {idx_path}")
            (version,) = struct.unpack("<Q", f.read(8))
            if version != 1:
                raise ValueError(f"bad version {version} in {idx_path}")
            (code,) = struct.unpack("<B", f.read(1))
            self.dtype = np.dtype(_DTYPE_BY_CODE[code])
            (self.sequence_count,) = struct.unpack("<Q", f.read(8))
            (self.document_count,) = struct.unpack("<Q", f.read(8))
            offset = f.tell()
        self._mmap = np.memmap(idx_path, mode="r", order="C")
        buf = memoryview(self._mmap)
        n, d = self.sequence_count, self.document_count
        self.sequence_lengths = np.frombuffer(buf, dtype=np.int32, count=n, offset=offset)
        self.sequence_pointers = np.frombuffer(buf, dtype=np.int64, count=n, offset=offset + 4 * n)
        self.document_indices = np.frombuffer(buf, dtype=np.int64, count=d, offset=offset + 12 * n)
        self.sequence_modes = None
        if multimodal:
            self.sequence_modes = np.frombuffer(buf, dtype=np.int8, count=n, offset=offset + 12 * n + 8 * d)
        if n and int(self.sequence_lengths.shape[0]) != n:
            raise ValueError(f"truncated index {idx_path}")


class MMapIndexedDataset:
    """Read side (indexed_dataset.py:340-523).  `ds[i]` -> tokens of sequence i; `ds.get(i, offset, length)` -> a slice of
    it; `sequence_lengths`, `document_indices` as in the reference."""

    def __init__(self, path_prefix: str, multimodal: bool = False):
        self.path_prefix = path_prefix
        self.multimodal = multimodal
        self.index = _Index(get_idx_path(path_prefix), multimodal)
        self.bin = np.memmap(get_bin_path(path_prefix), mode="r", order="C")
        self._tokens = self.bin.view(self.index.dtype) if self.bin.size else np.zeros(0, dtype=self.index.dtype)

    def __len__(self) -> int:
        return self.index.sequence_count

    @property
    def dtype(self) -> np.dtype:
        return self.index.dtype

    @property
    def sequence_lengths(self) -> np.ndarray:
        return self.index.sequence_lengths

    @property
    def document_indices(self) -> np.ndarray:
        return self.index.document_indices

    @property
    def sequence_element_offsets(self) -> np.ndarray:
        """element (not byte) offset of every sequence inside the token array"""
        return self.index.sequence_pointers // self.index.dtype.itemsize

    @property
    def tokens(self) -> np.ndarray:
        """the whole .bin as one 1-D array of token ids (memory mapped)"""
        return self._tokens

    def get(self, idx: int, offset: int = 0, length: int | None = None) -> np.ndarray:
        n = int(self.index.sequence_lengths[idx])
        if length is None:
            length = n - offset
        start = int(self.index.sequence_pointers[idx]) // self.index.dtype.itemsize + offset
        return self._tokens[start : start + length]

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            return self.get(int(idx))
        if isinstance(idx, slice):
            start, stop, step = idx.indices(len(self))
            if step != 1:
                raise ValueError("slices into the indexed dataset must be contiguous")
            lengths = self.index.sequence_lengths[idx]
            first = int(self.index.sequence_pointers[start]) // self.index.dtype.itemsize
            flat = self._tokens[first : first + int(lengths.sum())]
            return np.split(flat, np.cumsum(lengths)[:-1])
        raise TypeError(f"unexpected index type {type(idx)}")

    @staticmethod
    def exists(path_prefix: str) -> bool:
        return os.path.exists(get_idx_path(path_prefix)) and os.path.exists(get_bin_path(path_prefix))


class MMapIndexedDatasetBuilder:
    """Write side (indexed_dataset.py:525-609): add_item / end_document / finalize -> .bin + .idx"""

    def __init__(self, bin_path: str, dtype=np.int32, multimodal: bool = False):
        self._file = open(bin_path, "wb")
        self.dtype = np.dtype(dtype)
        self.multimodal = multimodal
        self.sequence_lengths: list[int] = []
        self.document_indices: list[int] = [0]
        self.sequence_modes: list[int] | None = [] if multimodal else None

    def add_item(self, tokens, mode: int = 0) -> None:
        arr = np.asarray(tokens).astype(self.dtype, copy=False)
        self._file.write(arr.tobytes(order="C"))
        self.sequence_lengths.append(int(arr.size))
        if self.multimodal:
            self.sequence_modes.append(mode)

    def add_document(self, tokens, lengths: list[int], modes: list[int] | None = None) -> None:
        arr = np.asarray(tokens).astype(self.dtype, copy=False)
        self._file.write(arr.tobytes(order="C"))
        self.sequence_lengths.extend(int(x) for x in lengths)
        self.document_indices.append(len(self.sequence_lengths))
        if self.multimodal:
            self.sequence_modes.extend(modes if modes is not None else [0] * len(lengths))

    def end_document(self) -> None:
        self.document_indices.append(len(self.sequence_lengths))

    def add_index(self, path_prefix: str) -> None:
        """append a whole existing store (indexed_dataset.py:579-598; tools/megatron_dataset/merge_data.py): its sequences
        and document boundaries follow the ones already written; the token bytes are streamed, not parsed"""
        other = _Index(get_idx_path(path_prefix), multimodal=self.multimodal)
        if other.dtype != self.dtype:
            raise ValueError(f"cannot merge {path_prefix} ({other.dtype}) into a {self.dtype} store")
        base = len(self.sequence_lengths)
        self.sequence_lengths.extend(int(x) for x in other.sequence_lengths)
        self.document_indices.extend(int(base + d) for d in other.document_indices[1:])
        if self.multimodal:
            self.sequence_modes.extend(int(m) for m in other.sequence_modes)
        with open(get_bin_path(path_prefix), "rb") as f:
            shutil.copyfileobj(f, self._file, 16 << 20)

    def finalize(self, idx_path: str) -> None:
        self._file.close()
        lengths = np.asarray(self.sequence_lengths, dtype=np.int32)
        pointers = np.zeros(len(lengths), dtype=np.int64)
        if len(lengths) > 1:
            np.cumsum(lengths[:-1].astype(np.int64) * self.dtype.itemsize, out=pointers[1:])
        with open(idx_path, "wb") as f:
            f.write(_INDEX_HEADER)
            f.write(struct.pack("<Q", 1))
            f.write(struct.pack("<B", _CODE_BY_DTYPE[self.dtype]))
            f.write(struct.pack("<Q", len(lengths)))
            f.write(struct.pack("<Q", len(self.document_indices)))
            f.write(lengths.tobytes(order="C"))
            f.write(pointers.tobytes(order="C"))
            f.write(np.asarray(self.document_indices, dtype=np.int64).tobytes(order="C"))
            if self.sequence_modes is not None:
                f.write(np.asarray(self.sequence_modes, dtype=np.int8).tobytes(order="C"))
