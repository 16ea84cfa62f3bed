/* dolomite-b200 data feed: host-side C ABI (no CUDA, no Python types) of lib/libdolomite_data.so.
 *
 * Drop-in for the pybind11 module the reference JIT-compiles at start-up (data/megatron/utils/helpers.cpp, loaded by
 * data/megatron/utils/__init__.py:21-35 `compile_helpers`) plus the batch assembly that the reference does sample by
 * sample in Python (data/megatron/gpt_dataset.py:117-160).  Plain pointers and sizes; the caller owns every buffer.
 * Python binding: dolomite_engine_b200/data/gpt_dataset.py (`ctypes`).  All functions are thread-safe and re-entrant.
 */
#ifndef DOLOMITE_DATA_H_
#define DOLOMITE_DATA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* (num_epochs * tokens_per_epoch - 1) / seq_length -- the sample count of helpers.cpp:94 / :172 */
int64_t dolomite_data_num_samples(int64_t seq_length, int64_t num_epochs, int64_t tokens_per_epoch);

/* Sample index of a GPT dataset -- replaces helpers.cpp:72-148 `build_sample_idx_int32` and :150-222
 * `build_sample_idx_int64` (dispatch on the dtype of the document index, data/megatron/utils/__init__.py:52-66).
 *   sizes[d]      tokens of document d (int32, MMapIndexedDataset.sequence_lengths)
 *   doc_idx[i]    document id at slot i of the shuffled, epoch-repeated document stream, n_doc_idx entries
 *   out           [num_samples + 1][2] = (slot, token offset) of the first token of every sample; sample k is the
 *                 seq_length + 1 tokens starting at stream position k * seq_length
 * returns the number of rows written (num_samples + 1), or a negative value when doc_idx holds fewer tokens than
 * num_epochs * tokens_per_epoch claims (the reference reads out of bounds in that case). */
int64_t dolomite_data_build_sample_index_i32(const int32_t* sizes, const int32_t* doc_idx, int64_t n_doc_idx,
                                             int64_t seq_length, int64_t num_epochs, int64_t tokens_per_epoch,
                                             int32_t* out);
int64_t dolomite_data_build_sample_index_i64(const int32_t* sizes, const int64_t* doc_idx, int64_t n_doc_idx,
                                             int64_t seq_length, int64_t num_epochs, int64_t tokens_per_epoch,
                                             int64_t* out);

/* Blend of several datasets by weight -- replaces helpers.cpp:17-70 `build_blending_indices` (called from
 * data/megatron/blended_dataset.py:122-128): sample i is drawn from the dataset with the largest
 * weights[d] * max(i, 1) - drawn[d]; dataset_index int16[size], dataset_sample_index int64[size]. */
void dolomite_data_build_blending_indices(int16_t* dataset_index, int64_t* dataset_sample_index, const double* weights,
                                          int32_t num_datasets, int64_t size);

/* One micro-batch straight from the memory-mapped .bin into (pinned) host memory as int64 -- the work of
 * GPTDataset._query_document_sample_shuffle_indices (gpt_dataset.py:117-160: per-sample numpy slices + concatenate +
 * astype(int64)) and of the DataLoader collate, for all rows at once.
 *   bin, elem_bytes          token array and token width in bytes (1, 2, 4, 8; 1/2 byte ids are unsigned)
 *   part_ptr / part_len      element offset and element count of every document slice
 *   row_first_part[r..r+1]   the slices of row r (n_rows + 1 entries)
 *   out                      [n_rows][row_len] int64
 * returns 0, or -1 / -3 when the slices of a row do not add up to row_len, -2 for an unsupported elem_bytes. */
int32_t dolomite_data_gather_rows(const void* bin, int32_t elem_bytes, const int64_t* part_ptr, const int64_t* part_len,
                                  const int64_t* row_first_part, int64_t n_rows, int64_t row_len, int64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* DOLOMITE_DATA_H_ */
