"""Pretraining data feed: Megatron .bin/.idx token stores -> `{"text": int64[mbs, S+1]}`"""

from .gpt_dataset import (
    BlendedDataset,
    GPTDataset,
    MegatronBatchSampler,
    PackedBatchLoader,
    build_blending_indices,
    build_gpt_datasets,
    build_sample_index,
    get_train_val_test_samples,
)
from .indexed_dataset import MMapIndexedDataset, MMapIndexedDatasetBuilder, get_bin_path, get_idx_path, optimal_dtype
from .fim import FIMSpec, HFTokenizerCodec, apply_fim
