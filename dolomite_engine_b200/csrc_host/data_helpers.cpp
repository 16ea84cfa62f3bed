// Host-side index builders of the pretraining data feed (C ABI, no CUDA, no Python types).
//
// Replaces the reference's pybind11 module data/megatron/utils/helpers.cpp (build_sample_idx_int32/int64 :72-222,
// build_blending_indices :20-70) behind plain pointers.  Both are restated from their definition rather than from the
// reference loops:
//
//  * sample index: the documents of `doc_idx` laid end to end form one token stream; sample k is the S+1 tokens starting
//    at stream position k*S (consecutive samples overlap by one token).  Row k of the index is therefore the (document
//    slot, offset) that CONTAINS stream position k*S -- one forward merge of the running document end against k*S.
//    Zero-length documents contain no position and are skipped, exactly like the reference's inner loop.
//  * blending index: greedy largest-deficit assignment, sample i goes to the dataset d maximising
//    weight[d] * max(i, 1) - count[d] (first maximum wins).
#include <cstdint>
#include <cstring>
#include <vector>

#if defined(__GNUC__)
#define DOLO_EXPORT extern "C" __attribute__((visibility("default")))
#else
#define DOLO_EXPORT extern "C"
#endif

namespace {

template <typename DocT, typename OutT>
int64_t build_sample_index_impl(const int32_t* sizes, const DocT* doc_idx, int64_t n_doc_idx, int64_t seq_length,
                                int64_t num_epochs, int64_t tokens_per_epoch, OutT* out) {
    const int64_t num_samples = (num_epochs * tokens_per_epoch - 1) / seq_length;
    int64_t slot = 0;       // index into doc_idx of the document under the cursor
    int64_t slot_begin = 0; // stream position of that document's first token
    for (int64_t k = 0; k <= num_samples; ++k) {
        const int64_t pos = k * seq_length;
        // advance to the document containing `pos`
        while (slot < n_doc_idx) {
            const int64_t len = sizes[doc_idx[slot]];
            if (pos < slot_begin + len) break;
            slot_begin += len;
            ++slot;
        }
        if (slot >= n_doc_idx) return -(k + 1);  // the stream is shorter than the caller claimed
        out[2 * k] = static_cast<OutT>(slot);
        out[2 * k + 1] = static_cast<OutT>(pos - slot_begin);
    }
    return num_samples + 1;
}

}  // namespace

// rows written (= num_samples + 1) or a negative number if doc_idx runs out of tokens.  `out` holds 2 * rows values.
DOLO_EXPORT int64_t dolomite_data_build_sample_index_i32(const int32_t* sizes, const int32_t* doc_idx, int64_t n_doc_idx,
                                                         int64_t seq_length, int64_t num_epochs,
                                                         int64_t tokens_per_epoch, int32_t* out) {
    return build_sample_index_impl(sizes, doc_idx, n_doc_idx, seq_length, num_epochs, tokens_per_epoch, out);
}
DOLO_EXPORT int64_t dolomite_data_build_sample_index_i64(const int32_t* sizes, const int64_t* doc_idx, int64_t n_doc_idx,
                                                         int64_t seq_length, int64_t num_epochs,
                                                         int64_t tokens_per_epoch, int64_t* out) {
    return build_sample_index_impl(sizes, doc_idx, n_doc_idx, seq_length, num_epochs, tokens_per_epoch, out);
}

DOLO_EXPORT int64_t dolomite_data_num_samples(int64_t seq_length, int64_t num_epochs, int64_t tokens_per_epoch) {
    return (num_epochs * tokens_per_epoch - 1) / seq_length;
}

// dataset_index[i] (int16, the reference's dtype) / dataset_sample_index[i] (int64): which dataset sample i of the blend comes from and its
// running index inside that dataset.
DOLO_EXPORT void dolomite_data_build_blending_indices(int16_t* dataset_index, int64_t* dataset_sample_index,
                                                      const double* weights, int32_t num_datasets, int64_t size) {
    std::vector<int64_t> counts(static_cast<size_t>(num_datasets > 0 ? num_datasets : 1), 0);
    for (int64_t i = 0; i < size; ++i) {
        const double denom = i > 1 ? static_cast<double>(i) : 1.0;
        int32_t best = 0;
        double best_deficit = weights[0] * denom - static_cast<double>(counts[0]);
        for (int32_t d = 1; d < num_datasets; ++d) {
            const double deficit = weights[d] * denom - static_cast<double>(counts[d]);
            if (deficit > best_deficit) {
                best = d;
                best_deficit = deficit;
            }
        }
        dataset_index[i] = static_cast<int16_t>(best);
        dataset_sample_index[i] = counts[best];
        ++counts[best];
    }
}

// Gather one micro-batch: rows[r] = the S+1 tokens of sample r assembled from up to `max_parts` document slices.
// part_ptr / part_len describe every slice (element offsets into `bin`, element counts), row_first_part[r] .. [r+1] the
// slices of row r.  Token width `elem_bytes` in {1, 2, 4, 8}; output int64 (the wrapper's batch["text"] dtype).
DOLO_EXPORT int32_t dolomite_data_gather_rows(const void* bin, int32_t elem_bytes, const int64_t* part_ptr,
                                              const int64_t* part_len, const int64_t* row_first_part, int64_t n_rows,
                                              int64_t row_len, int64_t* out) {
    for (int64_t r = 0; r < n_rows; ++r) {
        int64_t* dst = out + r * row_len;
        int64_t filled = 0;
        for (int64_t p = row_first_part[r]; p < row_first_part[r + 1]; ++p) {
            const int64_t n = part_len[p];
            if (filled + n > row_len) return -1;
            const int64_t off = part_ptr[p];
            switch (elem_bytes) {
                case 1: { const uint8_t* s = static_cast<const uint8_t*>(bin) + off; for (int64_t i = 0; i < n; ++i) dst[filled + i] = s[i]; break; }
                case 2: { const uint16_t* s = static_cast<const uint16_t*>(bin) + off; for (int64_t i = 0; i < n; ++i) dst[filled + i] = s[i]; break; }
                case 4: { const int32_t* s = static_cast<const int32_t*>(bin) + off; for (int64_t i = 0; i < n; ++i) dst[filled + i] = s[i]; break; }
                case 8: { const int64_t* s = static_cast<const int64_t*>(bin) + off; for (int64_t i = 0; i < n; ++i) dst[filled + i] = s[i]; break; }
                default: return -2;
            }
            filled += n;
        }
        if (filled != row_len) return -3;
    }
    return 0;
}
