// Shared pieces of the packed var-len causal attention kernels (forward and backward).
//
// qkv slot layout (reference: attention/padding_free.py:79-116): each token row holds `n_groups` groups of
// (q_per_group + 2) head slots of head_dim bf16: [q_0 .. q_{g-1}, k, v].  mha: n_groups = n_head, g = 1;
// gqa: n_groups = n_kv, g = n_head / n_kv; mqa: n_groups = 1, g = n_head.
//
// A [128 x head_dim] tile of Q, K or V is staged in shared memory as "chunks" along head_dim so that each chunk is
// one TMA box with a hardware swizzle that tcgen05 understands:  64-wide chunks (128 B rows, SWIZZLE_128B) followed
// by one remainder chunk of 32 (SWIZZLE_64B) or 16 (SWIZZLE_32B) columns.  head_dim 80 = 64 + 16, 128 = 64 + 64.
// The same bytes serve as a K-major operand (rows = MMA M/N, head_dim = contraction) and as an MN-major operand
// (head_dim = MMA N, rows = contraction), only the descriptor differs.
#pragma once
#include "common.cuh"

namespace dolo {

constexpr int ATT_TILE = 128;  // query rows / key rows per tile

template <int HD>
struct HeadChunks {
    static_assert(HD % 16 == 0 && HD >= 16 && HD <= 128, "head_dim must be a multiple of 16 in [16, 128]");
    static constexpr int NC64 = HD / 64;
    static constexpr int REM = HD % 64;
    static_assert(REM == 0 || REM == 16 || REM == 32, "head_dim % 64 must be 0, 16 or 32");
    static constexpr int NCHUNK = NC64 + (REM ? 1 : 0);
    static constexpr int TILE_BYTES = ATT_TILE * HD * 2;
    __host__ __device__ static constexpr int width(int c) { return c < NC64 ? 64 : REM; }
    __host__ __device__ static constexpr int col(int c) { return c * 64; }               // first head_dim column of chunk
    __host__ __device__ static constexpr int offset(int c) { return c * ATT_TILE * 128; }  // byte offset inside the tile
};

// UMMA layout_type field for a chunk of `w` bf16 columns
__host__ __device__ constexpr uint32_t chunk_layout_type(int w) { return w == 64 ? 2u : (w == 32 ? 4u : 6u); }

// K-major view of a chunk (rows = M or N of the MMA, chunk columns = contraction), k16 = which 16-wide K step
__device__ __forceinline__ uint64_t chunk_desc_kmajor(uint32_t chunk_saddr, int w, int k16) {
    return umma_smem_desc(chunk_saddr + k16 * 32, 16, 16 * w, chunk_layout_type(w));
}
// MN-major view (chunk columns = N of the MMA, rows = contraction), k16 = which group of 16 rows
__device__ __forceinline__ uint64_t chunk_desc_mnmajor(uint32_t chunk_saddr, int w, int k16) {
    return umma_smem_desc(chunk_saddr + k16 * 32 * w, 16 * w, 16 * w, chunk_layout_type(w));
}

// Attention-probability dropout (attention/base.py:252 `attn_dropout`; flash_attn_varlen_func(dropout_p=...) at
// attention/padding_free.py:49-59): P_ij is kept with probability 1 - p and scaled by 1 / (1 - p) AFTER the softmax
// normaliser was taken over the undropped row.  The mask is a hash of (global query token, global key token, head) and the
// keys of the call site, so the backward kernels regenerate it.  threshold == 0 <=> no dropout.
struct AttnDropout {
    uint32_t threshold;
    float keep_scale;
    uint32_t key0, key1;
};
__device__ __forceinline__ float attn_drop_scale(const AttnDropout& d, uint32_t head_key, int q_tok, int k_tok) {
    return dropout_hash_qk(uint32_t(q_tok), uint32_t(k_tok), head_key) >= d.threshold ? d.keep_scale : 0.f;
}

// Locate the (document, q-tile) of a linear tile index by scanning cu_seqlens (B is small: a few docs per row).
struct TileLoc {
    int doc_start;  // first token of the document
    int doc_len;
    int tile;       // 128-row tile index inside the document
    bool valid;
};
__device__ __forceinline__ TileLoc locate_tile(const int32_t* __restrict__ cu, int n_docs, int ti) {
    TileLoc r{0, 0, 0, false};
    int acc = 0;
    for (int d = 0; d < n_docs; ++d) {
        const int s = cu[d], e = cu[d + 1];
        const int nt = (e - s + ATT_TILE - 1) / ATT_TILE;
        if (ti < acc + nt) {
            r.doc_start = s;
            r.doc_len = e - s;
            r.tile = ti - acc;
            r.valid = true;
            return r;
        }
        acc += nt;
    }
    return r;
}

}  // namespace dolo
