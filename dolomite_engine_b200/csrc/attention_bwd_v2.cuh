// Attention backward, pipelined variant for head_dim <= 80 (included by attention_bwd.cu).
//
// Same math and operand views as attn_bwd_kernel (v1), but the query tile is processed as two 64-query halves with
// double-buffered S^T / dP^T accumulators in TMEM, a double-buffered dS^T staging tile in shared memory and a
// separate warp group draining dQ, so that the tensor pipe (S^T/dP^T of step s+1, dV/dK/dQ of step s-1) runs
// concurrently with the softmax warps working on step s:
//
//   warp 0      TMA producer  (K_j,V_j once; Q_i,dO_i 2-stage ring)
//   warp 1      MMA issuer    A_s: S^T_s = K Q_h^T, dP^T_s = V dO_h^T   (128 keys x 64 queries, buffers s & 1)
//                             C_s: dV += P^T_s dO_h, dK += dS^T_s Q_h   (A operands from TMEM)
//                             after the second half: dQ_i = dS_i K_j    (128 queries, A = dS^T tile in smem, MN-major)
//   warps 4-7   softmax       one key row per thread: P^T, dS^T -> TMEM (bf16, aliasing the consumed S^T/dP^T columns)
//                             and dS^T -> smem
//   warps 2,3,8,9  dQ drain   TMEM -> red.global.add.v4.f32 into the fp32 dQ workspace
//
// TMEM columns: S^T[2] 0..127, dP^T[2] 128..255, dV 256.., dK 256+HD.., dQ 256+2HD..  (<= 496 for HD = 80).
#pragma once

template <int HD>
__global__ void __launch_bounds__(320, 1)
    attn_bwd_kernel_v2(const __grid_constant__ CUtensorMap tq64, const __grid_constant__ CUtensorMap tqR,
                       const __grid_constant__ CUtensorMap to64, const __grid_constant__ CUtensorMap toR,
                       const BwdParams p) {
    using CH = HeadChunks<HD>;
    static_assert(256 + 3 * HD <= 512, "v2 needs a private dQ accumulator (head_dim <= 80)");
    constexpr int TILE_BYTES = CH::TILE_BYTES;
    constexpr int HALF = 64;
    constexpr uint32_t ST_COL = 0, DP_COL = 128, DV_COL = 256, DK_COL = 256 + HD, DQ_COL = 256 + 2 * HD;
    constexpr int DS_BYTES = 2 * ATT_TILE * 128;  // 128 keys x 128 queries bf16 (two 64-query MN chunks)

    const TileLoc loc = locate_tile(p.cu_seqlens, p.n_docs, int(blockIdx.x));
    if (!loc.valid) return;
    const int group = blockIdx.y;
    const int j = loc.tile;
    const int n_q_tiles = (loc.doc_len + ATT_TILE - 1) / ATT_TILE;
    const int n_i = n_q_tiles - j;
    const int n_it = n_i * p.q_per_group;
    const int n_steps = 2 * n_it;
    const int k_col = (group * (p.q_per_group + 2) + p.q_per_group) * HD;
    const int v_col = k_col + HD;
    const int kv_row = loc.doc_start + j * ATT_TILE;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_align_1024(smem_raw);
    uint8_t* sK = smem;
    uint8_t* sV = sK + TILE_BYTES;
    uint8_t* sQ = sV + TILE_BYTES;             // [2]
    uint8_t* sDO = sQ + 2 * TILE_BYTES;        // [2]
    uint8_t* sDS = sDO + 2 * TILE_BYTES;       // [2] x DS_BYTES
    float* sLSE = reinterpret_cast<float*>(sDS + 2 * DS_BYTES);  // [2][128] (log2 units)
    float* sDelta = sLSE + 2 * ATT_TILE;                         // [2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sDelta + 2 * ATT_TILE);
    uint64_t* kv_full = bars;            // 1
    uint64_t* qdo_full = bars + 1;       // [2]
    uint64_t* qdo_empty = bars + 3;      // [2]
    uint64_t* sdp_full = bars + 5;       // [2]  S^T/dP^T of buffer b ready
    uint64_t* pds_ready = bars + 7;      // [2]  128 arrivals: P^T/dS^T of buffer b written
    uint64_t* dq_full = bars + 9;        // [2]  dQ of tile parity ready (also: dS smem buffer of that parity is free)
    uint64_t* dq_done = bars + 11;       // 128 arrivals: dQ accumulator drained
    uint64_t* dkv_full = bars + 12;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tq64);
        tma_prefetch_desc(&to64);
        mbar_init(kv_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&qdo_full[i], 1);
            mbar_init(&qdo_empty[i], 1);
            mbar_init(&sdp_full[i], 1);
            mbar_init(&pds_ready[i], 128);
            mbar_init(&dq_full[i], 1);
        }
        mbar_init(dq_done, 128);
        mbar_init(dkv_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto load_tile = [&](uint8_t* dst, uint64_t* bar, const CUtensorMap* m64, const CUtensorMap* mR, int col, int row) {
#pragma unroll
        for (int c = 0; c < CH::NCHUNK; ++c) {
            const CUtensorMap* m = (c < CH::NC64) ? m64 : mR;
            tma_load_2d(dst + CH::offset(c), m, bar, col + CH::col(c), row);
        }
    };

    if (warp == 0) {
        // ======================= TMA producer =======================
        if (elect_one()) {
            mbar_expect_tx(kv_full, 2 * TILE_BYTES);
            load_tile(sK, kv_full, &tq64, &tqR, k_col, kv_row);
            load_tile(sV, kv_full, &tq64, &tqR, v_col, kv_row);
            for (int it = 0; it < n_it; ++it) {
                const int stage = it & 1;
                const uint32_t phase = uint32_t(it >> 1) & 1;
                const int s_head = it / n_i, i = j + (it - s_head * n_i);
                const int head = group * p.q_per_group + s_head;
                const int q_col = (group * (p.q_per_group + 2) + s_head) * HD;
                const int q_row = loc.doc_start + i * ATT_TILE;
                mbar_wait(&qdo_empty[stage], phase ^ 1, 30);
                mbar_expect_tx(&qdo_full[stage], 2 * TILE_BYTES);
                load_tile(sQ + stage * TILE_BYTES, &qdo_full[stage], &tq64, &tqR, q_col, q_row);
                load_tile(sDO + stage * TILE_BYTES, &qdo_full[stage], &to64, &toR, head * HD, q_row);
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        if (elect_one()) {
            constexpr uint32_t idesc_s = umma_idesc_bf16(128, HALF, false, false);
            mbar_wait(kv_full, 0, 31);
            const uint32_t k_s = smem_u32(sK), v_s = smem_u32(sV);

            auto issue_A = [&](int s) {
                const int it = s >> 1, h = s & 1, stage = it & 1, b = s & 1;
                if (h == 0) mbar_wait(&qdo_full[stage], uint32_t(it >> 1) & 1, 32);
                tc_fence_after();
                const uint32_t q_s = smem_u32(sQ + stage * TILE_BYTES);
                const uint32_t do_s = smem_u32(sDO + stage * TILE_BYTES);
                bool first = true;
#pragma unroll
                for (int c = 0; c < CH::NCHUNK; ++c) {
                    const int w = CH::width(c);
#pragma unroll
                    for (int k = 0; k < w / 16; ++k) {
                        umma_ss(tmem_base + ST_COL + b * HALF, chunk_desc_kmajor(k_s + CH::offset(c), w, k),
                                chunk_desc_kmajor(q_s + CH::offset(c) + h * HALF * 2 * w, w, k), idesc_s, first ? 0u : 1u);
                        first = false;
                    }
                }
                first = true;
#pragma unroll
                for (int c = 0; c < CH::NCHUNK; ++c) {
                    const int w = CH::width(c);
#pragma unroll
                    for (int k = 0; k < w / 16; ++k) {
                        umma_ss(tmem_base + DP_COL + b * HALF, chunk_desc_kmajor(v_s + CH::offset(c), w, k),
                                chunk_desc_kmajor(do_s + CH::offset(c) + h * HALF * 2 * w, w, k), idesc_s, first ? 0u : 1u);
                        first = false;
                    }
                }
                umma_commit(&sdp_full[b]);
            };

            issue_A(0);
            if (n_steps > 1) issue_A(1);
            for (int s = 0; s < n_steps; ++s) {
                const int it = s >> 1, h = s & 1, stage = it & 1, b = s & 1;
                mbar_wait(&pds_ready[b], uint32_t(it) & 1, 33);
                tc_fence_after();
                const uint32_t q_s = smem_u32(sQ + stage * TILE_BYTES);
                const uint32_t do_s = smem_u32(sDO + stage * TILE_BYTES);
#pragma unroll
                for (int c = 0; c < CH::NCHUNK; ++c) {
                    const int w = CH::width(c);
                    const uint32_t idesc_ts = umma_idesc_bf16(128, w, false, true);
#pragma unroll
                    for (int k = 0; k < HALF / 16; ++k)  // dV += P^T dO  (contraction over the 64 queries of this half)
                        umma_ts(tmem_base + DV_COL + CH::col(c), tmem_base + ST_COL + b * HALF + k * 8,
                                chunk_desc_mnmajor(do_s + CH::offset(c) + h * HALF * 2 * w, w, k), idesc_ts,
                                (s > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < HALF / 16; ++k)  // dK += dS^T Q
                        umma_ts(tmem_base + DK_COL + CH::col(c), tmem_base + DP_COL + b * HALF + k * 8,
                                chunk_desc_mnmajor(q_s + CH::offset(c) + h * HALF * 2 * w, w, k), idesc_ts,
                                (s > 0 || k > 0) ? 1u : 0u);
                }
                if (h == 1) {
                    if (it > 0) {
                        mbar_wait(dq_done, uint32_t(it - 1) & 1, 34);  // dQ accumulator drained
                        tc_fence_after();
                    }
                    const uint32_t ds_s = smem_u32(sDS + (it & 1) * DS_BYTES);
#pragma unroll
                    for (int c = 0; c < CH::NCHUNK; ++c) {
                        const int w = CH::width(c);
                        const uint32_t idesc_dq = umma_idesc_bf16(128, w, true, true);
#pragma unroll
                        for (int k = 0; k < ATT_TILE / 16; ++k)
                            umma_ss(tmem_base + DQ_COL + CH::col(c), umma_smem_desc(ds_s + k * 2048, ATT_TILE * 128, 1024, 2),
                                    chunk_desc_mnmajor(k_s + CH::offset(c), w, k), idesc_dq, k > 0 ? 1u : 0u);
                    }
                    umma_commit(&dq_full[it & 1]);
                    umma_commit(&qdo_empty[stage]);
                }
                if (s + 2 < n_steps) issue_A(s + 2);
                if (s == n_steps - 1) umma_commit(dkv_full);
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ======================= softmax warps: one key row per thread =======================
        const int sub = warp & 3;
        const int r = sub * 32 + lane;
        const uint32_t t_lane = tmem_base + (uint32_t(sub * 32) << 16);
        const int kj = j * ATT_TILE + r;
        const bool key_ok = kj < loc.doc_len;
        const float LOG2E = 1.4426950408889634f;
        for (int it = 0; it < n_it; ++it) {
            const int s_head = it / n_i, i = j + (it - s_head * n_i);
            const int head = group * p.q_per_group + s_head;
            float* lse_s = sLSE + (it & 1) * ATT_TILE;
            float* del_s = sDelta + (it & 1) * ATT_TILE;
            {
                const int qi = i * ATT_TILE + r;
                float l = INFINITY, d = 0.f;
                if (qi < loc.doc_len) {
                    const int64_t idx = int64_t(head) * p.T + loc.doc_start + qi;
                    l = p.lse[idx] * LOG2E;
                    d = p.delta[idx];
                }
                lse_s[r] = l;
                del_s[r] = d;
            }
            named_bar_sync(2, 128);
            if (it >= 2) mbar_wait(&dq_full[it & 1], uint32_t((it >> 1) - 1) & 1, 35);  // dS smem buffer free again
            const bool diag = (i == j);
            uint8_t* ds_buf = sDS + (it & 1) * DS_BYTES;
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const int b = h;
                mbar_wait(&sdp_full[b], uint32_t(it) & 1, 36);
                tc_fence_after();
#pragma unroll 1
                for (int ch = 0; ch < 2; ++ch) {
                    uint32_t sv[32], dv[32];
                    tmem_ld32(t_lane + ST_COL + b * HALF + ch * 32, sv);
                    tmem_ld32(t_lane + DP_COL + b * HALF + ch * 32, dv);
                    tmem_ld_wait();
                    uint32_t pp[16], dd[16];
#pragma unroll
                    for (int c2 = 0; c2 < 32; c2 += 2) {
                        float pv[2], dsv[2];
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int c = h * HALF + ch * 32 + c2 + u;  // query column inside the 128-query tile
                            float pe = fast_exp2(__uint_as_float(sv[c2 + u]) * p.scale_log2 - lse_s[c]);
                            if (!key_ok || (diag && r > c)) pe = 0.f;
                            pv[u] = pe;
                            dsv[u] = pe * (__uint_as_float(dv[c2 + u]) - del_s[c]) * p.scale;
                        }
                        pp[c2 >> 1] = pack_bf16(pv[0], pv[1]);
                        dd[c2 >> 1] = pack_bf16(dsv[0], dsv[1]);
                    }
                    tmem_st16(t_lane + ST_COL + b * HALF + ch * 16, pp);
                    tmem_st16(t_lane + DP_COL + b * HALF + ch * 16, dd);
                    uint8_t* rowp = ds_buf + h * (ATT_TILE * 128) + r * 128;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int piece = ch * 4 + q;
                        uint4 v = make_uint4(dd[q * 4], dd[q * 4 + 1], dd[q * 4 + 2], dd[q * 4 + 3]);
                        *reinterpret_cast<uint4*>(rowp + ((piece ^ (r & 7)) << 4)) = v;
                    }
                }
                tmem_st_wait();
                tc_fence_before();
                fence_proxy_async_smem();
                mbar_arrive(&pds_ready[b]);
            }
        }
        // ---------------- dK_j, dV_j epilogue ----------------
        mbar_wait(dkv_full, 0, 37);
        tc_fence_after();
        __nv_bfloat16* krow = p.dqkv + int64_t(kv_row + r) * p.row_stride + k_col;
        __nv_bfloat16* vrow = p.dqkv + int64_t(kv_row + r) * p.row_stride + v_col;
#pragma unroll 1
        for (int c0 = 0; c0 < HD; c0 += 16) {
            uint32_t a[16], bb[16];
            tmem_ld16(t_lane + DK_COL + c0, a);
            tmem_ld16(t_lane + DV_COL + c0, bb);
            tmem_ld_wait();
            if (key_ok) {
                uint4 x, y;
                x.x = pack_bf16(__uint_as_float(a[0]), __uint_as_float(a[1]));
                x.y = pack_bf16(__uint_as_float(a[2]), __uint_as_float(a[3]));
                x.z = pack_bf16(__uint_as_float(a[4]), __uint_as_float(a[5]));
                x.w = pack_bf16(__uint_as_float(a[6]), __uint_as_float(a[7]));
                y.x = pack_bf16(__uint_as_float(a[8]), __uint_as_float(a[9]));
                y.y = pack_bf16(__uint_as_float(a[10]), __uint_as_float(a[11]));
                y.z = pack_bf16(__uint_as_float(a[12]), __uint_as_float(a[13]));
                y.w = pack_bf16(__uint_as_float(a[14]), __uint_as_float(a[15]));
                *reinterpret_cast<uint4*>(krow + c0) = x;
                *reinterpret_cast<uint4*>(krow + c0 + 8) = y;
                x.x = pack_bf16(__uint_as_float(bb[0]), __uint_as_float(bb[1]));
                x.y = pack_bf16(__uint_as_float(bb[2]), __uint_as_float(bb[3]));
                x.z = pack_bf16(__uint_as_float(bb[4]), __uint_as_float(bb[5]));
                x.w = pack_bf16(__uint_as_float(bb[6]), __uint_as_float(bb[7]));
                y.x = pack_bf16(__uint_as_float(bb[8]), __uint_as_float(bb[9]));
                y.y = pack_bf16(__uint_as_float(bb[10]), __uint_as_float(bb[11]));
                y.z = pack_bf16(__uint_as_float(bb[12]), __uint_as_float(bb[13]));
                y.w = pack_bf16(__uint_as_float(bb[14]), __uint_as_float(bb[15]));
                *reinterpret_cast<uint4*>(vrow + c0) = x;
                *reinterpret_cast<uint4*>(vrow + c0 + 8) = y;
            }
        }
    } else {
        // ======================= dQ drain warps (2, 3, 8, 9): TMEM lane == query row =======================
        const int sub = warp & 3;
        const int r = sub * 32 + lane;
        const uint32_t t_lane = tmem_base + (uint32_t(sub * 32) << 16);
        for (int it = 0; it < n_it; ++it) {
            const int s_head = it / n_i, i = j + (it - s_head * n_i);
            const int head = group * p.q_per_group + s_head;
            mbar_wait(&dq_full[it & 1], uint32_t(it >> 1) & 1, 38);
            tc_fence_after();
            const int qi = i * ATT_TILE + r;
            const bool q_ok = qi < loc.doc_len;
            float* dst = p.dq_accum + (int64_t(head) * p.T + loc.doc_start + qi) * HD;
#pragma unroll 1
            for (int c0 = 0; c0 < HD; c0 += 16) {
                uint32_t o[16];
                tmem_ld16(t_lane + DQ_COL + c0, o);
                tmem_ld_wait();
                if (q_ok) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        red_add_v4(dst + c0 + q * 4, __uint_as_float(o[q * 4]), __uint_as_float(o[q * 4 + 1]),
                                   __uint_as_float(o[q * 4 + 2]), __uint_as_float(o[q * 4 + 3]));
                }
            }
            tc_fence_before();
            mbar_arrive(dq_done);
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
int launch_bwd_v2(const void* dout, const void* qkv, int64_t row_stride, const BwdParams& p, cudaStream_t st) {
    using CH = HeadChunks<HD>;
    CUtensorMap tq64, tqR, to64, toR;
    int rc = make_maps<HD>(qkv, row_stride, p.T, &tq64, &tqR);
    if (rc) return rc;
    rc = make_maps<HD>(dout, int64_t(p.n_heads) * HD, p.T, &to64, &toR);
    if (rc) return rc;
    constexpr int smem_bytes = 1024 + 6 * CH::TILE_BYTES + 2 * (2 * ATT_TILE * 128) + 4 * ATT_TILE * 4 + 160;
    static_assert(smem_bytes <= 232448, "attention backward v2 shared memory budget exceeded");
    auto kern = attn_bwd_kernel_v2<HD>;
    static bool attr_set = false;
    if (!attr_set) {
        DOLO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        attr_set = true;
    }
    const int64_t max_tiles = (p.T + ATT_TILE - 1) / ATT_TILE + p.n_docs;
    dim3 grid((unsigned)max_tiles, (unsigned)p.n_groups);
    kern<<<grid, 320, smem_bytes, st>>>(tq64, tqR, to64, toR, p);
    DOLO_LAUNCH_OK("attn_varlen_bwd_v2");
    return DOLO_OK;
}
