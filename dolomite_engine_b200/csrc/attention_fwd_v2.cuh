// Attention forward, ping-pong variant (included by attention_fwd.cu).
//
// One CTA owns TWO consecutive 128-row query tiles (A = 2t, B = 2t+1) of one head of one document and shares the K/V
// tiles between them.  Two softmax warp groups (one per query tile) alternate with the tensor pipe:
//
//   MMA warp :  S_A(0) S_B(0) | wait P_A(j): PV_A(j), S_A(j+1) | wait P_B(j): PV_B(j), S_B(j+1) | ...
//
// so while group A runs its softmax on S_A(j+1) the tensor core executes PV_B(j) and S_B(j+1), and vice versa: the MMAs
// are hidden behind the softmax instead of being serialised with it (v1 relies on two co-resident CTAs for overlap).
// TMEM: S_A 0..127, S_B 128..255, O_A 256.., O_B 256+HD..  (HD = 80: 416 columns; HD = 128: 512).
#pragma once

struct TilePairLoc {
    int doc_start, doc_len, pair;
    bool valid;
};
__device__ __forceinline__ TilePairLoc locate_tile_pair(const int32_t* __restrict__ cu, int n_docs, int pi) {
    TilePairLoc r{0, 0, 0, false};
    int acc = 0;
    for (int d = 0; d < n_docs; ++d) {
        const int s = cu[d], e = cu[d + 1];
        const int np = ((e - s + ATT_TILE - 1) / ATT_TILE + 1) / 2;
        if (pi < acc + np) {
            r.doc_start = s;
            r.doc_len = e - s;
            r.pair = pi - acc;
            r.valid = true;
            return r;
        }
        acc += np;
    }
    return r;
}

template <int HD>
__global__ void __launch_bounds__(320, 1)  // 10 warps = 3 per SM sub-partition -> at most 168 registers per thread
    attn_fwd_kernel_v2(const __grid_constant__ CUtensorMap tmap64, const __grid_constant__ CUtensorMap tmapR,
                       const FwdParams p) {
    using CH = HeadChunks<HD>;
    constexpr int TILE_BYTES = CH::TILE_BYTES;
    constexpr uint32_t O_COL = 256;

    const int pi = int(gridDim.x) - 1 - int(blockIdx.x);  // long (late) pairs first
    const TilePairLoc loc = locate_tile_pair(p.cu_seqlens, p.n_docs, pi);
    if (!loc.valid) return;
    const int head = blockIdx.y;
    const int group = head / p.q_per_group, slot = head % p.q_per_group;
    const int q_col = (group * (p.q_per_group + 2) + slot) * HD;
    const int k_col = (group * (p.q_per_group + 2) + p.q_per_group) * HD;
    const int v_col = k_col + HD;
    const int n_tiles_doc = (loc.doc_len + ATT_TILE - 1) / ATT_TILE;
    const int tile_a = 2 * loc.pair;
    const bool has_b = tile_a + 1 < n_tiles_doc;
    const int n_kv_a = tile_a + 1;                       // query tile A sees key tiles 0 .. tile_a
    const int n_kv = has_b ? tile_a + 2 : tile_a + 1;    // tile B sees one more
    const int row_base_a = loc.doc_start + tile_a * ATT_TILE;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_align_1024(smem_raw);
    uint8_t* sQ = smem;                         // [2] query tiles A, B
    uint8_t* sK = sQ + 2 * TILE_BYTES;          // [2] stages
    uint8_t* sV = sK + 2 * TILE_BYTES;          // [2] stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * TILE_BYTES);
    uint64_t* q_full = bars;          // 1
    uint64_t* kv_full = bars + 1;     // [2]
    uint64_t* kv_empty = bars + 3;    // [2]
    uint64_t* s_full = bars + 5;      // [2] per query tile
    uint64_t* p_ready = bars + 7;     // [2] per query tile, 128 arrivals
    uint64_t* o_full = bars + 9;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        if (CH::NC64 > 0) tma_prefetch_desc(&tmap64);
        if (CH::REM > 0) tma_prefetch_desc(&tmapR);
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&kv_empty[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&p_ready[i], 128);
            mbar_init(&o_full[i], 1);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto load_tile = [&](uint8_t* dst, uint64_t* bar, int col, int row) {
#pragma unroll
        for (int c = 0; c < CH::NCHUNK; ++c) {
            const CUtensorMap* m = (c < CH::NC64) ? &tmap64 : &tmapR;
            tma_load_2d(dst + CH::offset(c), m, bar, col + CH::col(c), row);
        }
    };

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(q_full, (has_b ? 2 : 1) * TILE_BYTES);
            load_tile(sQ, q_full, q_col, row_base_a);
            if (has_b) load_tile(sQ + TILE_BYTES, q_full, q_col, row_base_a + ATT_TILE);
            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j < n_kv; ++j) {
                mbar_wait(&kv_empty[stage], phase ^ 1, 40);
                mbar_expect_tx(&kv_full[stage], 2 * TILE_BYTES);
                const int krow = loc.doc_start + j * ATT_TILE;
                load_tile(sK + stage * TILE_BYTES, &kv_full[stage], k_col, krow);
                load_tile(sV + stage * TILE_BYTES, &kv_full[stage], v_col, krow);
                if (++stage == 2) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, false, false);
            mbar_wait(q_full, 0, 41);
            // S_x(j) = Q_x K_j^T into TMEM columns [128x, 128x+128)
            auto issue_S = [&](int x, int j) {
                const int stage = j & 1;
                const uint32_t q_s = smem_u32(sQ + x * TILE_BYTES);
                const uint32_t k_s = smem_u32(sK + stage * TILE_BYTES);
                bool first = true;
#pragma unroll
                for (int c = 0; c < CH::NCHUNK; ++c) {
                    const int w = CH::width(c);
#pragma unroll
                    for (int k = 0; k < w / 16; ++k) {
                        umma_ss(tmem_base + x * 128, chunk_desc_kmajor(q_s + CH::offset(c), w, k),
                                chunk_desc_kmajor(k_s + CH::offset(c), w, k), idesc_qk, first ? 0u : 1u);
                        first = false;
                    }
                }
                umma_commit(&s_full[x]);
            };
            // O_x += P_x(j) V_j   (A = P from TMEM columns [128x, 128x+64), B = V_j MN-major)
            auto issue_PV = [&](int x, int j) {
                const int stage = j & 1;
                const uint32_t v_s = smem_u32(sV + stage * TILE_BYTES);
#pragma unroll
                for (int c = 0; c < CH::NCHUNK; ++c) {
                    const int w = CH::width(c);
                    const uint32_t idesc_pv = umma_idesc_bf16(128, w, false, true);
#pragma unroll
                    for (int k = 0; k < ATT_TILE / 16; ++k)
                        umma_ts(tmem_base + O_COL + x * HD + CH::col(c), tmem_base + x * 128 + k * 8,
                                chunk_desc_mnmajor(v_s + CH::offset(c), w, k), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
                }
            };
            mbar_wait(&kv_full[0], 0, 42);
            tc_fence_after();
            issue_S(0, 0);
            if (has_b) issue_S(1, 0);
            for (int j = 0; j < n_kv; ++j) {
                const int stage = j & 1;
                const bool a_active = j < n_kv_a;
                const bool next_ready_needed = j + 1 < n_kv;
                if (a_active) {
                    mbar_wait(&p_ready[0], uint32_t(j & 1), 43);
                    tc_fence_after();
                    issue_PV(0, j);
                    if (j == n_kv_a - 1) umma_commit(&o_full[0]);
                }
                if (next_ready_needed) {
                    mbar_wait(&kv_full[stage ^ 1], uint32_t((j + 1) >> 1) & 1, 44);
                    tc_fence_after();
                    if (j + 1 < n_kv_a) issue_S(0, j + 1);
                }
                if (has_b) {
                    mbar_wait(&p_ready[1], uint32_t(j & 1), 45);
                    tc_fence_after();
                    issue_PV(1, j);
                    if (j == n_kv - 1) umma_commit(&o_full[1]);
                }
                umma_commit(&kv_empty[stage]);  // K_j / V_j no longer needed once everything issued so far retires
                if (has_b && next_ready_needed) issue_S(1, j + 1);
            }
        }
    } else {
        // ---------------- softmax groups: warps 2-5 -> query tile A, warps 6-9 -> query tile B ----------------
        const int x = (warp - 2) >> 2;
        const int sub = warp & 3;
        const int r = sub * 32 + lane;
        const uint32_t t_lane = tmem_base + (uint32_t(sub * 32) << 16);
        const uint32_t s_col = x * 128, o_col = O_COL + x * HD;
        const int tile = tile_a + x;
        const int my_n_kv = tile + 1;
        const int q0 = tile * ATT_TILE;
        const int qi = q0 + r;
        const bool active = (x == 0) || has_b;
        float m_run = -INFINITY, l_run = 0.f;
        if (active) {
            for (int j = 0; j < my_n_kv; ++j) {
                mbar_wait(&s_full[x], uint32_t(j & 1), 46);
                tc_fence_after();
                const bool diag = (j == my_n_kv - 1);
                const int kbase = j * ATT_TILE;
                float mx = m_run;
                float lsum = 0.f, lsum1 = 0.f;
                float alpha;
                // O *= alpha (previous PV has completed: s_full is committed after it); skipped when no row of this warp
                // raised its running max, the common case after the first tiles
                auto rescale_o = [&]() {
                    if (j > 0 && !__all_sync(0xffffffffu, alpha == 1.f)) {
#pragma unroll 1
                        for (int c0 = 0; c0 < HD; c0 += 16) {
                            uint32_t o[16];
                            tmem_ld16(t_lane + o_col + c0, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            tmem_st16(t_lane + o_col + c0, o);
                        }
                    }
                };
                if (!diag) {
                    // interior tile, single pass: the whole 128-score row lives in registers (one CTA per SM leaves ~200
                    // registers per thread), so TMEM is read once; row max on FMNMX3, scale/shift on FFMA2, sums on FADD2
                    uint32_t v0[32], v1[32], v2[32], v3[32];
                    tmem_ld32(t_lane + s_col, v0);
                    tmem_ld32(t_lane + s_col + 32, v1);
                    tmem_ld32(t_lane + s_col + 64, v2);
                    tmem_ld32(t_lane + s_col + 96, v3);
                    tmem_ld_wait();
                    reg_fence32(v0);
                    reg_fence32(v1);
                    reg_fence32(v2);
                    reg_fence32(v3);
                    float ma = mx, mb = -INFINITY, mc = -INFINITY, md = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        ma = fmax3(ma, __uint_as_float(v0[i]), __uint_as_float(v0[i + 1]));
                        mb = fmax3(mb, __uint_as_float(v1[i]), __uint_as_float(v1[i + 1]));
                        mc = fmax3(mc, __uint_as_float(v2[i]), __uint_as_float(v2[i + 1]));
                        md = fmax3(md, __uint_as_float(v3[i]), __uint_as_float(v3[i + 1]));
                    }
                    mx = fmaxf(fmax3(ma, mb, mc), md);
                    const float m_scaled = (mx == -INFINITY) ? 0.f : mx * p.scale_log2;
                    alpha = (m_run == -INFINITY) ? 0.f : fast_exp2(m_run * p.scale_log2 - m_scaled);
                    rescale_o();
                    const float neg_m = -m_scaled;
                    auto expo = [&](const uint32_t (&v)[32], int ch) {
                        uint32_t pk[16];
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            float x0, x1;
                            ffma2_bcast(x0, x1, __uint_as_float(v[i]), __uint_as_float(v[i + 1]), p.scale_log2, neg_m);
                            const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
                            fadd2(lsum, lsum1, p0, p1);
                            pk[i >> 1] = pack_bf16(p0, p1);
                        }
                        tmem_st16(t_lane + s_col + ch * 16, pk);  // bf16 P over the S columns (all of S is in registers)
                    };
                    expo(v0, 0);
                    expo(v1, 1);
                    expo(v2, 2);
                    expo(v3, 3);
                } else {
#pragma unroll 1
                    for (int ch = 0; ch < 4; ++ch) {
                        uint32_t v[32];
                        tmem_ld32(t_lane + s_col + ch * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            float s = __uint_as_float(v[i]);
                            if (kbase + ch * 32 + i > qi) s = -INFINITY;
                            mx = fmaxf(mx, s);
                        }
                    }
                    const float m_scaled = (mx == -INFINITY) ? 0.f : mx * p.scale_log2;
                    alpha = (m_run == -INFINITY) ? 0.f : fast_exp2(m_run * p.scale_log2 - m_scaled);
                    rescale_o();
                    const float neg_m = -m_scaled;
#pragma unroll 1
                    for (int ch = 0; ch < 4; ++ch) {
                        uint32_t v[32];
                        tmem_ld32(t_lane + s_col + ch * 32, v);
                        tmem_ld_wait();
                        uint32_t pk[16];
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            float p0 = fast_exp2(fmaf(__uint_as_float(v[i]), p.scale_log2, neg_m));
                            float p1 = fast_exp2(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, neg_m));
                            if (kbase + ch * 32 + i > qi) p0 = 0.f;
                            if (kbase + ch * 32 + i + 1 > qi) p1 = 0.f;
                            lsum += p0;
                            lsum1 += p1;
                            pk[i >> 1] = pack_bf16(p0, p1);
                        }
                        tmem_st16(t_lane + s_col + ch * 16, pk);
                    }
                }
                const float m_new = mx;
                lsum += lsum1;
                l_run = l_run * alpha + lsum;
                m_run = m_new;
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&p_ready[x]);
            }
            // ---------------- epilogue ----------------
            mbar_wait(&o_full[x], 0, 47);
            tc_fence_after();
            const bool row_ok = qi < loc.doc_len;
            const float inv_l = l_run > 0.f ? 1.f / l_run : 0.f;
            const int64_t grow = int64_t(loc.doc_start) + qi;
            __nv_bfloat16* orow = p.out + grow * (int64_t(p.n_heads) * HD) + int64_t(head) * HD;
#pragma unroll 1
            for (int c0 = 0; c0 < HD; c0 += 16) {
                uint32_t o[16];
                tmem_ld16(t_lane + o_col + c0, o);
                tmem_ld_wait();
                if (row_ok) {
                    uint4 a, b;
                    a.x = pack_bf16(__uint_as_float(o[0]) * inv_l, __uint_as_float(o[1]) * inv_l);
                    a.y = pack_bf16(__uint_as_float(o[2]) * inv_l, __uint_as_float(o[3]) * inv_l);
                    a.z = pack_bf16(__uint_as_float(o[4]) * inv_l, __uint_as_float(o[5]) * inv_l);
                    a.w = pack_bf16(__uint_as_float(o[6]) * inv_l, __uint_as_float(o[7]) * inv_l);
                    b.x = pack_bf16(__uint_as_float(o[8]) * inv_l, __uint_as_float(o[9]) * inv_l);
                    b.y = pack_bf16(__uint_as_float(o[10]) * inv_l, __uint_as_float(o[11]) * inv_l);
                    b.z = pack_bf16(__uint_as_float(o[12]) * inv_l, __uint_as_float(o[13]) * inv_l);
                    b.w = pack_bf16(__uint_as_float(o[14]) * inv_l, __uint_as_float(o[15]) * inv_l);
                    *reinterpret_cast<uint4*>(orow + c0) = a;
                    *reinterpret_cast<uint4*>(orow + c0 + 8) = b;
                }
            }
            if (row_ok) p.lse[int64_t(head) * p.T + grow] = m_run * p.scale + logf(l_run);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

template <int HD>
int launch_fwd_v2(const void* qkv, int64_t row_stride, const FwdParams& p, cudaStream_t st) {
    using CH = HeadChunks<HD>;
    CUtensorMap t64, tR;
    uint64_t dims[2] = {uint64_t(row_stride), uint64_t(p.T)};
    uint64_t strides[2] = {2, uint64_t(row_stride) * 2};
    uint32_t box[2] = {64, ATT_TILE};
    int rc;
    if (CH::NC64 > 0) {
        rc = dolo_make_tmap(&t64, qkv, 2, 2, dims, strides, box, DOLO_SW_128);
        if (rc) return rc;
    }
    if (CH::REM > 0) {
        box[0] = CH::REM;
        rc = dolo_make_tmap(&tR, qkv, 2, 2, dims, strides, box, CH::REM == 32 ? DOLO_SW_64 : DOLO_SW_32);
        if (rc) return rc;
    }
    if (CH::NC64 == 0) t64 = tR;
    if (CH::REM == 0) tR = t64;
    constexpr int smem_bytes = 1024 + 6 * CH::TILE_BYTES + 160;
    static_assert(smem_bytes <= 232448, "attention forward v2 shared memory budget exceeded");
    auto kern = attn_fwd_kernel_v2<HD>;
    static bool attr_set = false;
    if (!attr_set) {
        DOLO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        attr_set = true;
    }
    // upper bound on the number of query-tile pairs without reading cu_seqlens on the host
    const int64_t max_pairs = ((p.T + ATT_TILE - 1) / ATT_TILE + p.n_docs + 1) / 2 + p.n_docs;
    dim3 grid((unsigned)max_pairs, (unsigned)p.n_heads);
    kern<<<grid, 320, smem_bytes, st>>>(t64, tR, p);
    DOLO_LAUNCH_OK("attn_varlen_fwd_v2");
    return DOLO_OK;
}
