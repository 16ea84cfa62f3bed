// Training-mode dropout on flat bf16 activations (reference: nn.Dropout at gpt_dolomite/base.py:138 `drop` after the
// embeddings, attention/base.py:92 + padding_free.py:75 `resid_dropout` after the attention c_proj, gpt_dolomite/mlp.py:43-49
// after the MLP c_proj, moe_dolomite/moe/base.py:106-120 after the expert combine), fused with what follows it in the block
// (gpt_dolomite/layer.py:73-86: `* m_residual`, `+ residual`; gpt_dolomite/base.py:368-371: `* m_emb`).
//
//   forward :  y = bf16(x * s)            s = 1 / (1 - p) where the element is kept, 0 where it is dropped
//              y = bf16(y * post_mul)     (only when post_mul != 1: m_residual / m_emb are separate bf16 multiplies)
//              y = bf16(residual + y)     (only with a residual)
//   backward:  g = bf16(dy * pre_mul)     (only when pre_mul != 1)
//              dx = bf16(g * s)
// The mask is a pure function of (element index, key0, key1) -- common.cuh: dropout_hash_flat -- so backward and
// recomputed (checkpointed) blocks regenerate it instead of storing it.  HBM-bound: 4 (6 with a residual) bytes / element.
#include "common.cuh"
#include "../../include/dolomite_b200.h"

using namespace dolo;

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
    f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 v;
    v.x = pack_bf16(f[0], f[1]); v.y = pack_bf16(f[2], f[3]); v.z = pack_bf16(f[4], f[5]); v.w = pack_bf16(f[6], f[7]);
    return v;
}

template <bool HAS_RES>
__global__ void __launch_bounds__(kThreads)
    dropout_fwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ res, uint4* __restrict__ out, int64_t n8,
                       uint32_t threshold, float keep_scale, float post_mul, uint32_t key0, uint32_t key1) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += int64_t(gridDim.x) * blockDim.x) {
        float f[8], r[8];
        unpack8(x[i], f);
        if (HAS_RES) unpack8(res[i], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool keep = dropout_hash_flat(uint64_t(i) * 8 + j, key0, key1) >= threshold;
            float y = bf16_round(f[j] * (keep ? keep_scale : 0.f));
            if (post_mul != 1.f) y = bf16_round(y * post_mul);
            if (HAS_RES) y = r[j] + y;
            f[j] = y;
        }
        out[i] = pack8(f);
    }
}

__global__ void __launch_bounds__(kThreads)
    dropout_bwd_kernel(const uint4* __restrict__ dy, uint4* __restrict__ dx, int64_t n8, uint32_t threshold, float keep_scale,
                       float pre_mul, uint32_t key0, uint32_t key1) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += int64_t(gridDim.x) * blockDim.x) {
        float f[8];
        unpack8(dy[i], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool keep = dropout_hash_flat(uint64_t(i) * 8 + j, key0, key1) >= threshold;
            float g = f[j];
            if (pre_mul != 1.f) g = bf16_round(g * pre_mul);
            f[j] = g * (keep ? keep_scale : 0.f);
        }
        dx[i] = pack8(f);
    }
}

inline int grid_for(int64_t work, int threads) {
    int64_t g = (work + threads - 1) / threads;
    const int64_t cap = int64_t(dolo_num_sms()) * 8;
    return int(g < 1 ? 1 : (g > cap ? cap : g));
}
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

uint32_t dolo_dropout_threshold(float p) {
    double t = double(p) * 4294967296.0;
    if (t < 0) t = 0;
    if (t > 4294967295.0) t = 4294967295.0;
    return uint32_t(t + 0.5 > 4294967295.0 ? 4294967295.0 : t + 0.5);
}

extern "C" int dolomite_b200_dropout_fwd(const void* x, const void* residual, void* out, int64_t n, float p, float post_mul,
                                         uint32_t key0, uint32_t key1, void* stream) {
    DOLO_REQUIRE(p >= 0.f && p < 1.f, "dropout: p=%f must be in [0, 1)", double(p));
    DOLO_REQUIRE(n >= 0 && n % 8 == 0, "dropout: n=%lld must be a multiple of 8", (long long)n);
    DOLO_REQUIRE(aligned16(x) && aligned16(out) && aligned16(residual), "dropout: pointers must be 16-byte aligned");
    if (n == 0) return DOLO_OK;
    const uint32_t thr = dolo_dropout_threshold(p);
    const float ks = 1.f / (1.f - p);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (residual != nullptr)
        dropout_fwd_kernel<true><<<grid_for(n / 8, kThreads), kThreads, 0, st>>>(
            static_cast<const uint4*>(x), static_cast<const uint4*>(residual), static_cast<uint4*>(out), n / 8, thr, ks,
            post_mul, key0, key1);
    else
        dropout_fwd_kernel<false><<<grid_for(n / 8, kThreads), kThreads, 0, st>>>(
            static_cast<const uint4*>(x), nullptr, static_cast<uint4*>(out), n / 8, thr, ks, post_mul, key0, key1);
    DOLO_LAUNCH_OK("dropout_fwd");
    return DOLO_OK;
}

extern "C" int dolomite_b200_dropout_bwd(const void* dy, void* dx, int64_t n, float p, float pre_mul, uint32_t key0,
                                         uint32_t key1, void* stream) {
    DOLO_REQUIRE(p >= 0.f && p < 1.f, "dropout: p=%f must be in [0, 1)", double(p));
    DOLO_REQUIRE(n >= 0 && n % 8 == 0, "dropout: n=%lld must be a multiple of 8", (long long)n);
    DOLO_REQUIRE(aligned16(dy) && aligned16(dx), "dropout: pointers must be 16-byte aligned");
    if (n == 0) return DOLO_OK;
    dropout_bwd_kernel<<<grid_for(n / 8, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(dy), static_cast<uint4*>(dx), n / 8, dolo_dropout_threshold(p), 1.f / (1.f - p), pre_mul,
        key0, key1);
    DOLO_LAUNCH_OK("dropout_bwd");
    return DOLO_OK;
}
