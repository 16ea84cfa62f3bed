// Attention backward, pipelined variant v3 for head_dim <= 80 (included by attention_bwd.cu).
// v3 = v2 with TWO softmax warp groups (2 warps per SM sub-partition instead of 1: each group owns 32 of the 64 query
// columns of a half step), mask-free fast path for interior tiles, vectorised LSE/Delta reads, dK / dV epilogue split
// between the groups.  ncu on v2 (profiles/r01_ncu_prof_attn_bwd_v2.txt): issue-slot 27 %, tensor 19 % -> latency bound
// on one softmax warp per scheduler.
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
//   warps 4-11  softmax       two groups x (one key row per thread, 32 query columns each): P^T, dS^T -> TMEM (bf16,
//                             aliasing the consumed S^T/dP^T columns) and dS^T -> smem
//   warps 2,3,12,13  dQ drain TMEM -> red.global.add.v4.f32 into the fp32 dQ workspace
//
// TMEM columns: S^T[2] 0..127, dP^T[2] 128..255, dV 256.., dK 256+HD.., dQ 256+2HD..  (<= 496 for HD = 80).
#pragma once

// K, V, 2 x (Q, dO), 2 x dS^T staging, dQ staging (4 warps x 2 swizzled 4 KB slabs), LSE/Delta (2 x 2 x 128 floats), barriers
template <int HD>
constexpr int bwd_v3_smem_need() {
    return 6 * HeadChunks<HD>::TILE_BYTES + 2 * (2 * ATT_TILE * 128) + 4 * 2 * 4096 + 4096 + 160;
}

// Uniform 16-column chunks (32-byte rows, SWIZZLE_32B) for head dims that are not a multiple of 64.  Every chunk is one K = 16
// step of the K-major views (S^T, dP^T), and the MN-major views (dV, dK, dQ: head_dim = MMA N) span all chunks through ONE
// descriptor (leading-dimension byte offset = chunk stride), so those products take one N = head_dim MMA per K step instead of
// one per chunk.  Why it matters (ncu call 77 + the clock64 timeline of call 79): the kernel is bound by shared-memory
// bandwidth (tensor-core operand reads 1952 + LSU 1565 wavefronts per tile of ~4700 cycles); with the 64 + 16 split every
// dQ K step read the 4 KB dS^T slice twice.
template <int HD>
struct UniformChunks {
    static constexpr int NC64 = 0;
    static constexpr int REM = 16;
    static constexpr int NCHUNK = HD / 16;
    static constexpr int TILE_BYTES = ATT_TILE * HD * 2;
    static constexpr int CHUNK_BYTES = ATT_TILE * 32;
    __host__ __device__ static constexpr int width(int) { return 16; }
    __host__ __device__ static constexpr int col(int c) { return c * 16; }
    __host__ __device__ static constexpr int offset(int c) { return c * CHUNK_BYTES; }
};
template <int HD, bool UNI>
struct BwdChunks {
    using type = HeadChunks<HD>;
};
template <int HD>
struct BwdChunks<HD, true> {
    using type = UniformChunks<HD>;
};

// NG = number of softmax warp groups (4 warps each): every group owns 64 / NG query columns of a half step.  NG = 4
// (704 threads, <= 88 registers) halves the dependent instruction chain per warp and doubles the warps each scheduler can
// interleave; the softmax warps, not the tensor pipe, bound the NG = 2 version (ncu: tensor pipe 20 %, issue 25 %).
template <int HD, int NG, bool LEAN, bool UNI>
__global__ void __launch_bounds__(32 * (6 + 4 * NG), 1)
    attn_bwd_kernel_v3(const __grid_constant__ CUtensorMap tq64, const __grid_constant__ CUtensorMap tqR,
                       const __grid_constant__ CUtensorMap to64, const __grid_constant__ CUtensorMap toR,
                       const __grid_constant__ CUtensorMap tdq32, const __grid_constant__ CUtensorMap tdq16,
                       const BwdParams p) {
    using CH = typename BwdChunks<HD, UNI>::type;
    static_assert(256 + 3 * HD <= 512, "v2 needs a private dQ accumulator (head_dim <= 80)");
    constexpr int TILE_BYTES = CH::TILE_BYTES;
    constexpr int HALF = 64;
    constexpr int COLS = HALF / NG;       // query columns per softmax warp per half step
    constexpr int NSC = COLS / 16;        // 16-column sub-chunks per warp
    constexpr int SOFTMAX_THREADS = 128 * NG;
    constexpr int FIRST_TAIL_WARP = 4 + 4 * NG;  // two more dQ drain warps after the softmax warps
    static_assert(NG == 2 || NG == 4, "two or four softmax groups");
    constexpr uint32_t ST_COL = 0, DP_COL = 128, DV_COL = 256, DK_COL = 256 + HD, DQ_COL = 256 + 2 * HD;
    constexpr int DS_BYTES = 2 * ATT_TILE * 128;  // 128 keys x 128 queries bf16 (two 64-query MN chunks)

    int cta_tile, group;  // key tiles of a document in natural order = longest first
    attn_cta_order(p.head_chunk, p.n_tile_slots, cta_tile, group);
    const TileLoc loc = locate_tile(p.cu_seqlens, p.n_docs, cta_tile);
    if (!loc.valid) return;
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
    uint8_t* sDQ = sDS + 2 * DS_BYTES;         // 4 drain warps x 2 slabs x 4 KB (32 rows x 32 fp32, 128B-swizzled; 1024-aligned)
    float* sLSE = reinterpret_cast<float*>(sDQ + 4 * 2 * 4096);  // [2][128] (log2 units)
    float* sDelta = sLSE + 2 * ATT_TILE;                         // [2][128]   (LEAN: the 4 KB hold per-warp private copies)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sLSE + 1024);
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
    // DEBUG timeline (dolomite_b200_debug_attn_bwd_trace): one thread per role of ONE CTA stamps clock64() at its events
    constexpr int TRACE_CAP = 2048;
    const bool tr_cta = p.trace != nullptr && cta_tile == p.trace_cta && group == 0;
    int tr_n = 0;
    auto TR = [&](int role, int id) {
        if (tr_cta && tr_n < TRACE_CAP)
            p.trace[role * TRACE_CAP + tr_n++] = (static_cast<unsigned long long>(id) << 56) | (clock64() & 0xFFFFFFFFFFFFFFull);
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tq64);
        tma_prefetch_desc(&to64);
        tma_prefetch_desc(&tdq32);
        mbar_init(kv_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&qdo_full[i], 1);
            mbar_init(&qdo_empty[i], 1);
            mbar_init(&sdp_full[i], 1);
            mbar_init(&pds_ready[i], SOFTMAX_THREADS);
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
            int s_head = 0, i = j - 1;  // (head slot, query tile) of iteration `it`, tracked without a division
            for (int it = 0; it < n_it; ++it) {
                const int stage = it & 1;
                const uint32_t phase = uint32_t(it >> 1) & 1;
                if (++i == n_q_tiles) {
                    i = j;
                    ++s_head;
                }
                const int head = group * p.q_per_group + s_head;
                const int q_col = (group * (p.q_per_group + 2) + s_head) * HD;
                const int q_row = loc.doc_start + i * ATT_TILE;
                mbar_wait(&qdo_empty[stage], phase ^ 1, 30);
                TR(0, 1);
                mbar_expect_tx(&qdo_full[stage], 2 * TILE_BYTES);
                load_tile(sQ + stage * TILE_BYTES, &qdo_full[stage], &tq64, &tqR, q_col, q_row);
                load_tile(sDO + stage * TILE_BYTES, &qdo_full[stage], &to64, &toR, head * HD, q_row);
                TR(0, 2);
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        if (elect_one()) {  // uniform single-thread region: no per-MMA R2UR waterfall loops
            constexpr uint32_t idesc_s = umma_idesc_bf16(128, HALF, false, false);
            mbar_wait(kv_full, 0, 31);
            const uint32_t k_s = smem_u32(sK), v_s = smem_u32(sV);

            auto issue_A = [&](int s) {
                const int it = s >> 1, h = s & 1, stage = it & 1, b = s & 1;
                if (h == 0) {
                    TR(1, 10);
                    mbar_wait(&qdo_full[stage], uint32_t(it >> 1) & 1, 32);
                    TR(1, 11);
                }
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
                TR(1, 12);
            };

            issue_A(0);
            if (n_steps > 1) issue_A(1);
            for (int s = 0; s < n_steps; ++s) {
                const int it = s >> 1, h = s & 1, stage = it & 1, b = s & 1;
                TR(1, 1);
                mbar_wait(&pds_ready[b], uint32_t(it) & 1, 33);
                TR(1, 2);
                tc_fence_after();
                const uint32_t q_s = smem_u32(sQ + stage * TILE_BYTES);
                const uint32_t do_s = smem_u32(sDO + stage * TILE_BYTES);
                if constexpr (UNI) {
                    constexpr uint32_t idesc_ts = umma_idesc_bf16(128, HD, false, true);
                    if (!(p.ablate & 16)) {
#pragma unroll
                        for (int k = 0; k < HALF / 16; ++k)  // dV += P^T dO: one N = HD MMA per 16 queries
                            umma_ts(tmem_base + DV_COL, tmem_base + ST_COL + b * HALF + ((16 * k) / COLS) * COLS + (((16 * k) % COLS) / 16) * 8,
                                    umma_smem_desc(do_s + h * HALF * 32 + k * 512, CH::CHUNK_BYTES, 256, 6), idesc_ts,
                                    (s > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                        for (int k = 0; k < HALF / 16; ++k)  // dK += dS^T Q
                            umma_ts(tmem_base + DK_COL, tmem_base + DP_COL + b * HALF + ((16 * k) / COLS) * COLS + (((16 * k) % COLS) / 16) * 8,
                                    umma_smem_desc(q_s + h * HALF * 32 + k * 512, CH::CHUNK_BYTES, 256, 6), idesc_ts,
                                    (s > 0 || k > 0) ? 1u : 0u);
                    }
                } else if (!(p.ablate & 16))
#pragma unroll
                for (int c = 0; c < CH::NCHUNK; ++c) {
                    const int w = CH::width(c);
                    const uint32_t idesc_ts = umma_idesc_bf16(128, w, false, true);
#pragma unroll
                    for (int k = 0; k < HALF / 16; ++k)  // dV += P^T dO  (contraction over the 64 queries of this half)
                        umma_ts(tmem_base + DV_COL + CH::col(c), tmem_base + ST_COL + b * HALF + ((16 * k) / COLS) * COLS + (((16 * k) % COLS) / 16) * 8,
                                chunk_desc_mnmajor(do_s + CH::offset(c) + h * HALF * 2 * w, w, k), idesc_ts,
                                (s > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < HALF / 16; ++k)  // dK += dS^T Q
                        umma_ts(tmem_base + DK_COL + CH::col(c), tmem_base + DP_COL + b * HALF + ((16 * k) / COLS) * COLS + (((16 * k) % COLS) / 16) * 8,
                                chunk_desc_mnmajor(q_s + CH::offset(c) + h * HALF * 2 * w, w, k), idesc_ts,
                                (s > 0 || k > 0) ? 1u : 0u);
                }
                // S^T / dP^T of step s+2 go first: the softmax warps are the critical path and must never wait behind the
                // dQ drain (the dQ MMA below needs the previous tile's accumulator drained by the red.add warps)
                TR(1, 3);
                // the dV / dK products of the second half are the last readers of this tile's Q / dO stage: hand it back to the
                // TMA producer now (the clock64 timeline of call 79 showed the MMA thread waiting ~470 cycles per tile for the
                // next tile's Q / dO when the release was tied to the dQ product, which does not read them)
                if (h == 1) umma_commit(&qdo_empty[stage]);
                if (s + 2 < n_steps) issue_A(s + 2);
                if (h == 1) {
                    if (it > 0) {
                        TR(1, 4);
                        mbar_wait(dq_done, uint32_t(it - 1) & 1, 34);  // dQ accumulator drained
                        TR(1, 5);
                        tc_fence_after();
                    }
                    const uint32_t ds_s = smem_u32(sDS + (it & 1) * DS_BYTES);
                    if constexpr (UNI) {
                        constexpr uint32_t idesc_dq = umma_idesc_bf16(128, HD, true, true);
                        if (!(p.ablate & 4)) {
#pragma unroll
                            for (int k = 0; k < ATT_TILE / 16; ++k)
                                umma_ss(tmem_base + DQ_COL, umma_smem_desc(ds_s + k * 2048, ATT_TILE * 128, 1024, 2),
                                        umma_smem_desc(k_s + k * 512, CH::CHUNK_BYTES, 256, 6), idesc_dq, k > 0 ? 1u : 0u);
                        }
                    } else if (!(p.ablate & 4))
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
                    TR(1, 6);
                }
                if (s == n_steps - 1) umma_commit(dkv_full);
            }
        }
    } else if (LEAN && warp >= 4 && warp < FIRST_TAIL_WARP) {
        // ======================= softmax warps, lean instruction stream =======================
        // Same roles as the block below; what changed (ncu call 70: these warps are never waiting for the tensor pipe, they ARE
        // the critical path at ~0.16 IPC each, 9.5 warp instructions per element):
        //   * no integer division per tile (the (head, query tile) pair is tracked incrementally: the div's MUFU.RCP chain queued
        //     behind the other warp's EX2 burst and cost 10 % of the samples);
        //   * statistics are warp-private (-lse*log2e, -delta of this warp's columns; lanes fetch them one tile ahead): no
        //     8-warp named barrier per tile, the warps drift apart and overlap each other's MUFU and FMA phases;
        //   * packed FFMA2 / FADD2 / FMUL2 (two elements per issue slot) for the exponent argument and dS;
        //   * the softmax scale is NOT applied here: dS' = P o (dP - delta) goes to the tensor pipe, dK is scaled in the epilogue
        //     and dQ in attn_dq_finalize_kernel (one multiply per OUTPUT element instead of one per score);
        //   * the generic->async proxy fence for the dS^T shared-memory tile is only needed before the second half's arrive
        //     (the dQ MMA is the only async-proxy reader and is issued after both halves).
        const int wg = (warp - 4) >> 2;
        const int sub = warp & 3;
        const int r = sub * 32 + lane;
        const uint32_t t_lane = tmem_base + (uint32_t(sub * 32) << 16);
        const int kj = j * ATT_TILE + r;
        const bool key_ok = kj < loc.doc_len;
        const bool tile_full = (j + 1) * ATT_TILE <= loc.doc_len;
        const float LOG2E = 1.4426950408889634f;
        // [0, COLS) -lse*log2e half 0 | [COLS, 2 COLS) half 1 | [2 COLS, 3 COLS) -delta half 0 | [3 COLS, 4 COLS) half 1
        float* st = sLSE + (warp - 4) * (4 * COLS);
        float l0, l1, d0, d1;  // raw statistics of the NEXT tile (lane l < COLS: query column wg * COLS + l of both halves)
        auto fetch_raw = [&](int s_head, int i) {
            const int head = group * p.q_per_group + s_head;
            const int q0 = i * ATT_TILE + wg * COLS + lane;
            const int64_t off = int64_t(head) * p.T + loc.doc_start + q0;
            const bool ok0 = lane < COLS && q0 < loc.doc_len, ok1 = lane < COLS && q0 + HALF < loc.doc_len;
            l0 = ok0 ? p.lse[off] : INFINITY;  // +inf -> P = 0 for query rows past the document
            l1 = ok1 ? p.lse[off + HALF] : INFINITY;
            d0 = ok0 ? p.delta[off] : 0.f;
            d1 = ok1 ? p.delta[off + HALF] : 0.f;
        };
        int s_head = 0, i = j;
        fetch_raw(s_head, i);
        uint32_t sv0[16], dv0[16];
        mbar_wait(&sdp_full[0], 0, 36);
        tc_fence_after();
        tmem_ld16(t_lane + ST_COL + wg * COLS, sv0);
        tmem_ld16(t_lane + DP_COL + wg * COLS, dv0);

        const int tr_role = (lane == 0 && sub == 0) ? 2 + wg : -1;
        auto TS = [&](int id) {
            if (tr_role >= 0) TR(tr_role, id);
        };
        for (int it = 0; it < n_it; ++it) {
            TS(1);
            __syncwarp();  // every lane is done with the previous tile's statistics
            if (lane < COLS) {
                st[lane] = -LOG2E * l0;
                st[COLS + lane] = -LOG2E * l1;
                st[2 * COLS + lane] = -d0;
                st[3 * COLS + lane] = -d1;
            }
            __syncwarp();
            const bool diag = (i == j);
            const bool need_mask = diag || !tile_full;
            if (++i == n_q_tiles) {
                i = j;
                ++s_head;
            }
            if (it + 1 < n_it) fetch_raw(s_head, i);
            if (it >= 2) mbar_wait(&dq_full[it & 1], uint32_t((it >> 1) - 1) & 1, 35);  // dS smem buffer free again
            TS(2);
            uint8_t* ds_buf = sDS + (it & 1) * DS_BYTES;

            auto process16 = [&](const uint32_t (&sv)[16], const uint32_t (&dv)[16], int h, int b, int sc) {
                const float* nls = st + h * COLS + sc * 16;
                const float* nds = st + 2 * COLS + h * COLS + sc * 16;
                uint32_t pp[8], dd[8];
                if (p.ablate & 8) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        pp[q] = sv[q] & 0x3f803f80u;
                        dd[q] = dv[q] & 0x3f803f80u;
                    }
                } else if (need_mask) {
                    const int cbase = h * HALF + wg * COLS + sc * 16;  // first query column (inside the 128-query tile)
#pragma unroll
                    for (int c2 = 0; c2 < 16; c2 += 2) {
                        float pv[2], dsv[2];
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            float pe = fast_exp2(fmaf(__uint_as_float(sv[c2 + u]), p.scale_log2, nls[c2 + u]));
                            if (!key_ok || (diag && r > cbase + c2 + u)) pe = 0.f;
                            pv[u] = pe;
                            dsv[u] = pe * (__uint_as_float(dv[c2 + u]) + nds[c2 + u]);
                        }
                        pp[c2 >> 1] = pack_bf16(pv[0], pv[1]);
                        dd[c2 >> 1] = pack_bf16(dsv[0], dsv[1]);
                    }
                } else {
#pragma unroll
                    for (int c4 = 0; c4 < 16; c4 += 4) {
                        const float4 l4 = *reinterpret_cast<const float4*>(nls + c4);
                        const float4 d4 = *reinterpret_cast<const float4*>(nds + c4);
                        float x0, x1, x2, x3, t0, t1, t2, t3, e0, e1, e2, e3;
                        ffma2_sv(x0, x1, __uint_as_float(sv[c4]), __uint_as_float(sv[c4 + 1]), p.scale_log2, l4.x, l4.y);
                        ffma2_sv(x2, x3, __uint_as_float(sv[c4 + 2]), __uint_as_float(sv[c4 + 3]), p.scale_log2, l4.z, l4.w);
                        const float p0 = fast_exp2(x0), p1 = fast_exp2(x1), p2 = fast_exp2(x2), p3 = fast_exp2(x3);
                        fadd2_v(t0, t1, __uint_as_float(dv[c4]), __uint_as_float(dv[c4 + 1]), d4.x, d4.y);
                        fadd2_v(t2, t3, __uint_as_float(dv[c4 + 2]), __uint_as_float(dv[c4 + 3]), d4.z, d4.w);
                        fmul2_v(e0, e1, p0, p1, t0, t1);
                        fmul2_v(e2, e3, p2, p3, t2, t3);
                        pp[c4 >> 1] = pack_bf16(p0, p1);
                        pp[(c4 >> 1) + 1] = pack_bf16(p2, p3);
                        dd[c4 >> 1] = pack_bf16(e0, e1);
                        dd[(c4 >> 1) + 1] = pack_bf16(e2, e3);
                    }
                }
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};" ::"r"(pp[0]),
                    "r"(pp[1]), "r"(pp[2]), "r"(pp[3]), "r"(pp[4]), "r"(pp[5]), "r"(pp[6]), "r"(pp[7]),
                    "r"(t_lane + ST_COL + b * HALF + wg * COLS + sc * 8)
                    : "memory");
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};" ::"r"(dd[0]),
                    "r"(dd[1]), "r"(dd[2]), "r"(dd[3]), "r"(dd[4]), "r"(dd[5]), "r"(dd[6]), "r"(dd[7]),
                    "r"(t_lane + DP_COL + b * HALF + wg * COLS + sc * 8)
                    : "memory");
                uint8_t* rowp = ds_buf + h * (ATT_TILE * 128) + r * 128;
                if (!(p.ablate & 32))
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int piece = wg * (COLS / 8) + sc * 2 + q;
                    *reinterpret_cast<uint4*>(rowp + ((piece ^ (r & 7)) << 4)) =
                        make_uint4(dd[q * 4], dd[q * 4 + 1], dd[q * 4 + 2], dd[q * 4 + 3]);
                }
            };

#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const int b = h;
                const uint32_t s_addr = t_lane + ST_COL + b * HALF + wg * COLS;
                const uint32_t d_addr = t_lane + DP_COL + b * HALF + wg * COLS;
                TS(3);
                tmem_ld_wait();
                TS(4);
                reg_fence16(sv0);
                reg_fence16(dv0);
                if constexpr (NSC == 2) {
                    uint32_t sv1[16], dv1[16];
                    tmem_ld16(s_addr + 16, sv1);
                    tmem_ld16(d_addr + 16, dv1);
                    process16(sv0, dv0, h, b, 0);
                    TS(5);
                    tmem_ld_wait();
                    TS(6);
                    reg_fence16(sv1);
                    reg_fence16(dv1);
                    process16(sv1, dv1, h, b, 1);
                } else {
                    process16(sv0, dv0, h, b, 0);
                }
                TS(7);
                const int s_next = 2 * it + h + 1;
                if (s_next < 2 * n_it) {
                    const int nb = s_next & 1;
                    mbar_wait(&sdp_full[nb], uint32_t(s_next >> 1) & 1, 36);
                    TS(8);
                    tc_fence_after();
                    tmem_ld16(t_lane + ST_COL + nb * HALF + wg * COLS, sv0);
                    tmem_ld16(t_lane + DP_COL + nb * HALF + wg * COLS, dv0);
                }
                tmem_st_wait();
                TS(9);
                tc_fence_before();
                if (h == 1) fence_proxy_async_smem();
                mbar_arrive(&pds_ready[b]);
                TS(10);
            }
        }
        // ---------------- epilogue: even groups store dK_j (scaled here), odd groups dV_j ----------------
        mbar_wait(dkv_full, 0, 37);
        tc_fence_after();
        {
            const bool is_dk = (wg & 1) == 0;
            const float osc = is_dk ? p.scale : 1.f;
            const uint32_t src_col = is_dk ? DK_COL : DV_COL;
            __nv_bfloat16* drow = p.dqkv + int64_t(kv_row + r) * p.row_stride + (is_dk ? k_col : v_col);
            constexpr int NCH = HD / 16, SPLIT = (NCH + 1) / 2;
            const int ch0 = (NG == 2) ? 0 : ((wg >> 1) == 0 ? 0 : SPLIT);
            const int ch1 = (NG == 2) ? NCH : ((wg >> 1) == 0 ? SPLIT : NCH);
#pragma unroll 1
            for (int c0 = ch0 * 16; c0 < ch1 * 16; c0 += 16) {
                uint32_t a[16];
                tmem_ld16(t_lane + src_col + c0, a);
                tmem_ld_wait();
                if (key_ok) {
                    uint4 x, y;
                    x.x = pack_bf16(__uint_as_float(a[0]) * osc, __uint_as_float(a[1]) * osc);
                    x.y = pack_bf16(__uint_as_float(a[2]) * osc, __uint_as_float(a[3]) * osc);
                    x.z = pack_bf16(__uint_as_float(a[4]) * osc, __uint_as_float(a[5]) * osc);
                    x.w = pack_bf16(__uint_as_float(a[6]) * osc, __uint_as_float(a[7]) * osc);
                    y.x = pack_bf16(__uint_as_float(a[8]) * osc, __uint_as_float(a[9]) * osc);
                    y.y = pack_bf16(__uint_as_float(a[10]) * osc, __uint_as_float(a[11]) * osc);
                    y.z = pack_bf16(__uint_as_float(a[12]) * osc, __uint_as_float(a[13]) * osc);
                    y.w = pack_bf16(__uint_as_float(a[14]) * osc, __uint_as_float(a[15]) * osc);
                    *reinterpret_cast<uint4*>(drow + c0) = x;
                    *reinterpret_cast<uint4*>(drow + c0 + 8) = y;
                }
            }
        }
    } else if (warp >= 4 && warp < FIRST_TAIL_WARP) {
        // ======================= softmax warps: 2 groups x (one key row per thread, 32 query columns) =======================
        const int wg = (warp - 4) >> 2;  // column group: query columns [32*wg, 32*wg+32) of each 64-query half
        const int sub = warp & 3;
        const int r = sub * 32 + lane;
        const int tidsm = (warp - 4) * 32 + lane;
        const uint32_t t_lane = tmem_base + (uint32_t(sub * 32) << 16);
        const int kj = j * ATT_TILE + r;
        const bool key_ok = kj < loc.doc_len;
        const bool tile_full = (j + 1) * ATT_TILE <= loc.doc_len;  // every key row of this CTA is inside the document
        const float LOG2E = 1.4426950408889634f;
        // LSE (log2 units) / Delta of a query tile: thread t < 128 fetches lse[t], thread 128 + t fetches delta[t].  The
        // fetch for tile it+1 is issued while tile it is processed (global latency off the critical path).
        auto fetch_stat = [&](int it) -> float {
            const int s_head = it / n_i, i = j + (it - s_head * n_i);
            const int head = group * p.q_per_group + s_head;
            const int qi = i * ATT_TILE + (tidsm & (ATT_TILE - 1));
            const bool q_ok = qi < loc.doc_len;
            const int64_t off = int64_t(head) * p.T + loc.doc_start + qi;
            if (tidsm < ATT_TILE) return q_ok ? p.lse[off] : INFINITY;  // scaled to log2 units when stored (not here:
                                                                        // the multiply would wait for the load at once)
            if (tidsm < 2 * ATT_TILE) return q_ok ? p.delta[off] : 0.f;
            return 0.f;
        };
        float stat_next = fetch_stat(0);
        uint32_t sv0[16], dv0[16];  // S^T / dP^T values of the NEXT half step's first sub-chunk (requested one step ahead)
        mbar_wait(&sdp_full[0], 0, 36);
        tc_fence_after();
        tmem_ld16(t_lane + ST_COL + wg * COLS, sv0);
        tmem_ld16(t_lane + DP_COL + wg * COLS, dv0);

        for (int it = 0; it < n_it; ++it) {
            const int s_head = it / n_i, i = j + (it - s_head * n_i);
            float* lse_s = sLSE + (it & 1) * ATT_TILE;
            float* del_s = sDelta + (it & 1) * ATT_TILE;
            if (tidsm < ATT_TILE) lse_s[tidsm] = stat_next * LOG2E;
            else if (tidsm < 2 * ATT_TILE) del_s[tidsm - ATT_TILE] = stat_next;
            named_bar_sync(2, SOFTMAX_THREADS);
            if (it + 1 < n_it) stat_next = fetch_stat(it + 1);
            if (it >= 2) mbar_wait(&dq_full[it & 1], uint32_t((it >> 1) - 1) & 1, 35);  // dS smem buffer free again
            const bool drop = p.drop.threshold != 0;  // dropout: every tile takes the per-element path
            const bool need_mask = (i == j) || !tile_full || drop;
            const bool diag = (i == j);
            const int q_tok0 = loc.doc_start + i * ATT_TILE;  // global token of query column 0 of this tile
            const uint32_t head_key = dropout_head_key(uint32_t(group * p.q_per_group + s_head), p.drop.key0, p.drop.key1);
            uint8_t* ds_buf = sDS + (it & 1) * DS_BYTES;

            // one 16-column sub-chunk: P^T = exp2(S^T*scale - lse), dS^T = scale * P^T o (dP^T - delta) -> bf16 pairs,
            // written to TMEM (A operands of dV / dK) and, for dS^T, to the MN-major smem tile (A operand of dQ)
            auto process16 = [&](const uint32_t (&sv)[16], const uint32_t (&dv)[16], int h, int b, int sc) {
                const int cbase = h * HALF + wg * COLS + sc * 16;  // first query column (inside the 128-query tile)
                uint32_t pp[8], dd[8];
                if (need_mask) {
#pragma unroll
                    for (int c2 = 0; c2 < 16; c2 += 2) {
                        float pv[2], dsv[2];
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int c = cbase + c2 + u;
                            float pe = fast_exp2(__uint_as_float(sv[c2 + u]) * p.scale_log2 - lse_s[c]);
                            if (!key_ok || (diag && r > c)) pe = 0.f;
                            float dpe = __uint_as_float(dv[c2 + u]);
                            pv[u] = pe;
                            if (drop) {
                                const float ms = attn_drop_scale(p.drop, head_key, q_tok0 + c, loc.doc_start + kj);
                                pv[u] = pe * ms;
                                dpe *= ms;
                            }
                            dsv[u] = pe * (dpe - del_s[c]) * p.scale;
                        }
                        pp[c2 >> 1] = pack_bf16(pv[0], pv[1]);
                        dd[c2 >> 1] = pack_bf16(dsv[0], dsv[1]);
                    }
                } else {
#pragma unroll
                    for (int c4 = 0; c4 < 16; c4 += 4) {
                        const float4 l4 = *reinterpret_cast<const float4*>(lse_s + cbase + c4);
                        const float4 d4 = *reinterpret_cast<const float4*>(del_s + cbase + c4);
                        const float p0 = fast_exp2(__uint_as_float(sv[c4]) * p.scale_log2 - l4.x);
                        const float p1 = fast_exp2(__uint_as_float(sv[c4 + 1]) * p.scale_log2 - l4.y);
                        const float p2 = fast_exp2(__uint_as_float(sv[c4 + 2]) * p.scale_log2 - l4.z);
                        const float p3 = fast_exp2(__uint_as_float(sv[c4 + 3]) * p.scale_log2 - l4.w);
                        pp[c4 >> 1] = pack_bf16(p0, p1);
                        pp[(c4 >> 1) + 1] = pack_bf16(p2, p3);
                        dd[c4 >> 1] = pack_bf16(p0 * p.scale * (__uint_as_float(dv[c4]) - d4.x),
                                                p1 * p.scale * (__uint_as_float(dv[c4 + 1]) - d4.y));
                        dd[(c4 >> 1) + 1] = pack_bf16(p2 * p.scale * (__uint_as_float(dv[c4 + 2]) - d4.z),
                                                      p3 * p.scale * (__uint_as_float(dv[c4 + 3]) - d4.w));
                    }
                }
                // bf16 P^T / dS^T alias fp32 columns this thread has already consumed: group wg reads fp32 columns
                // [32wg + 16sc, +16) and writes bf16 columns [32wg + 8sc, +8) -- always inside its own consumed range
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};" ::"r"(pp[0]),
                    "r"(pp[1]), "r"(pp[2]), "r"(pp[3]), "r"(pp[4]), "r"(pp[5]), "r"(pp[6]), "r"(pp[7]),
                    "r"(t_lane + ST_COL + b * HALF + wg * COLS + sc * 8)
                    : "memory");
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};" ::"r"(dd[0]),
                    "r"(dd[1]), "r"(dd[2]), "r"(dd[3]), "r"(dd[4]), "r"(dd[5]), "r"(dd[6]), "r"(dd[7]),
                    "r"(t_lane + DP_COL + b * HALF + wg * COLS + sc * 8)
                    : "memory");
                uint8_t* rowp = ds_buf + h * (ATT_TILE * 128) + r * 128;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int piece = wg * (COLS / 8) + sc * 2 + q;
                    *reinterpret_cast<uint4*>(rowp + ((piece ^ (r & 7)) << 4)) =
                        make_uint4(dd[q * 4], dd[q * 4 + 1], dd[q * 4 + 2], dd[q * 4 + 3]);
                }
            };

#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const int b = h;
                // sv0 / dv0 of this half step were requested one half step ago (or before the loop): the TMEM read latency and
                // the wait for S^T / dP^T overlap the previous step's tcgen05.st drain, proxy fence and barrier arrive (ncu of the
                // unpipelined loop: 10 % of the samples sat on `fence.proxy.async` + `mbarrier.arrive` behind those stores)
                const uint32_t s_addr = t_lane + ST_COL + b * HALF + wg * COLS;
                const uint32_t d_addr = t_lane + DP_COL + b * HALF + wg * COLS;
                tmem_ld_wait();
                reg_fence16(sv0);
                reg_fence16(dv0);
                if constexpr (NSC == 2) {
                    uint32_t sv1[16], dv1[16];
                    tmem_ld16(s_addr + 16, sv1);
                    tmem_ld16(d_addr + 16, dv1);
                    process16(sv0, dv0, h, b, 0);
                    tmem_ld_wait();
                    reg_fence16(sv1);
                    reg_fence16(dv1);
                    process16(sv1, dv1, h, b, 1);
                } else {
                    process16(sv0, dv0, h, b, 0);
                }
                const int s_next = 2 * it + h + 1;
                if (s_next < 2 * n_it) {  // request the first sub-chunk of the next half step (the other S^T / dP^T buffer)
                    const int nb = s_next & 1;
                    mbar_wait(&sdp_full[nb], uint32_t(s_next >> 1) & 1, 36);
                    tc_fence_after();
                    tmem_ld16(t_lane + ST_COL + nb * HALF + wg * COLS, sv0);
                    tmem_ld16(t_lane + DP_COL + nb * HALF + wg * COLS, dv0);
                }
                tmem_st_wait();
                tc_fence_before();
                fence_proxy_async_smem();
                mbar_arrive(&pds_ready[b]);
            }
        }
        // ---------------- epilogue: group 0 stores dK_j, group 1 stores dV_j ----------------
        mbar_wait(dkv_full, 0, 37);
        tc_fence_after();
        {
            // even groups store dK_j, odd groups dV_j; with four groups each takes half of the 16-column chunks
            const bool is_dk = (wg & 1) == 0;
            const uint32_t src_col = is_dk ? DK_COL : DV_COL;
            __nv_bfloat16* drow = p.dqkv + int64_t(kv_row + r) * p.row_stride + (is_dk ? k_col : v_col);
            constexpr int NCH = HD / 16, SPLIT = (NCH + 1) / 2;
            const int ch0 = (NG == 2) ? 0 : ((wg >> 1) == 0 ? 0 : SPLIT);
            const int ch1 = (NG == 2) ? NCH : ((wg >> 1) == 0 ? SPLIT : NCH);
#pragma unroll 1
            for (int c0 = ch0 * 16; c0 < ch1 * 16; c0 += 16) {
                uint32_t a[16];
                tmem_ld16(t_lane + src_col + c0, a);
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
                    *reinterpret_cast<uint4*>(drow + c0) = x;
                    *reinterpret_cast<uint4*>(drow + c0 + 8) = y;
                }
            }
        }
    } else {
        // ======================= dQ drain warps (2, 3, 12, 13): TMEM lane == query row =======================
        // Each warp moves its 32 rows of the dQ accumulator through two private 4 KB slabs (32 rows x 32 fp32 columns, written
        // in the 128-byte swizzle the TMA unit expects, so the stores are bank-conflict free) and hands every slab to the TMA
        // engine as ONE tile reduce-add into dq_accum viewed as [heads * T, HD]; the L2 performs the adds.  History: per-thread
        // red.global.add.v4.f32 cost 26 % of the kernel (profiles/r01_probe_call18_dq_bulk_reduce.jsonl); the first slab version
        // was a linear [32][HD] block for a 1-D bulk reduce, whose 320-byte row pitch made every store a 4-way bank conflict
        // (ncu call 71: 33 M of the kernel's 83 M shared-memory wavefronts were store replays).  Rows past the document add
        // zeros (their dS is zero: lse = +inf), rows past the tensor are clipped by the TMA unit.
        const int sub = warp & 3;
        const uint32_t t_lane = tmem_base + (uint32_t(sub * 32) << 16);
        uint8_t* slab = sDQ + sub * (2 * 4096);
        int nb = 0;
        int s_head = 0, i = j - 1;
        for (int it = 0; it < n_it; ++it) {
            if (++i == n_q_tiles) {
                i = j;
                ++s_head;
            }
            const int head = group * p.q_per_group + s_head;
            const int row0 = int(int64_t(head) * p.T + loc.doc_start + i * ATT_TILE + sub * 32);
            if (lane == 0 && warp == 2) TR(4, 1);
            mbar_wait(&dq_full[it & 1], uint32_t(it >> 1) & 1, 38);
            if (lane == 0 && warp == 2) TR(4, 2);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 + 32 <= HD && !(p.ablate & 2); c0 += 32) {
                uint8_t* buf = slab + (nb & 1) * 4096;
                if (lane == 0) tma_store_wait_read<1>();  // the reduce issued from this slab two chunks ago has read it
                __syncwarp();
                uint32_t o[32];
                tmem_ld32(t_lane + DQ_COL + c0, o);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<uint4*>(buf + lane * 128 + ((q ^ (lane & 7)) << 4)) =
                        make_uint4(o[q * 4], o[q * 4 + 1], o[q * 4 + 2], o[q * 4 + 3]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0 && !(p.ablate & 1)) {
                    tma_reduce_add_2d(&tdq32, buf, c0, row0);
                    tma_store_commit();
                }
                ++nb;
            }
            if (HD % 32 == 16 && !(p.ablate & 2)) {  // last 16 columns: 64-byte rows, 64-byte swizzle
                uint8_t* buf = slab + (nb & 1) * 4096;
                if (lane == 0) tma_store_wait_read<1>();
                __syncwarp();
                uint32_t o[16];
                tmem_ld16(t_lane + DQ_COL + (HD - 16), o);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<uint4*>(buf + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) =
                        make_uint4(o[q * 4], o[q * 4 + 1], o[q * 4 + 2], o[q * 4 + 3]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0 && !(p.ablate & 1)) {
                    tma_reduce_add_2d(&tdq16, buf, HD - 16, row0);
                    tma_store_commit();
                }
                ++nb;
            }
            tc_fence_before();
            mbar_arrive(dq_done);  // accumulator free for the next dQ MMA
            if (lane == 0 && warp == 2) TR(4, 3);
        }
        if (lane == 0) tma_store_wait_all<0>();  // shared memory must outlive the last reduce
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

template <int HD, int NG, bool LEAN, bool UNI>
int launch_bwd_v3(const void* dout, const void* qkv, int64_t row_stride, const BwdParams& p, cudaStream_t st) {
    CUtensorMap tq64, tqR, to64, toR;
    int rc;
    if constexpr (UNI) {  // one map per tensor: boxes of 16 columns x 128 rows, 32-byte swizzle
        uint32_t box[2] = {16, ATT_TILE};
        uint64_t dims[2] = {uint64_t(row_stride), uint64_t(p.T)};
        uint64_t strides[2] = {2, uint64_t(row_stride) * 2};
        rc = dolo_make_tmap(&tqR, qkv, 2, 2, dims, strides, box, DOLO_SW_32);
        if (rc) return rc;
        dims[0] = uint64_t(p.n_heads) * HD;
        strides[1] = dims[0] * 2;
        rc = dolo_make_tmap(&toR, dout, 2, 2, dims, strides, box, DOLO_SW_32);
        if (rc) return rc;
        tq64 = tqR;
        to64 = toR;
    } else {
        rc = make_maps<HD>(qkv, row_stride, p.T, &tq64, &tqR);
        if (rc) return rc;
        rc = make_maps<HD>(dout, int64_t(p.n_heads) * HD, p.T, &to64, &toR);
        if (rc) return rc;
    }
    CUtensorMap tdq32, tdq16;
    {
        // dq_accum as [heads * T rows, HD columns] fp32; boxes = 32 rows x 32 (16) columns, the drain warps' slabs
        uint64_t dims[2] = {uint64_t(HD), uint64_t(p.n_heads) * uint64_t(p.T)};
        uint64_t strides[2] = {4, uint64_t(HD) * 4};
        uint32_t box[2] = {32, 32};
        rc = dolo_make_tmap(&tdq32, p.dq_accum, 4, 2, dims, strides, box, DOLO_SW_128);
        if (rc) return rc;
        box[0] = 16;
        rc = dolo_make_tmap(&tdq16, p.dq_accum, 4, 2, dims, strides, box, DOLO_SW_64);
        if (rc) return rc;
    }
    constexpr int need = bwd_v3_smem_need<HD>();
    constexpr int smem_bytes = need + 1024;
    static_assert(smem_bytes <= 232448, "attention backward v3 shared memory budget exceeded");
    auto kern = attn_bwd_kernel_v3<HD, NG, LEAN, UNI>;
    static bool attr_set = false;
    if (!attr_set) {
        DOLO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        attr_set = true;
    }
    const int64_t max_tiles = (p.T + ATT_TILE - 1) / ATT_TILE + p.n_docs;
    DOLO_REQUIRE(max_tiles == p.n_tile_slots && max_tiles * p.n_groups < (1ll << 31), "attn_bwd: grid too large");
    dim3 grid((unsigned)(max_tiles * p.n_groups));
    kern<<<grid, 32 * (6 + 4 * NG), smem_bytes, st>>>(tq64, tqR, to64, toR, tdq32, tdq16, p);
    DOLO_LAUNCH_OK("attn_varlen_bwd_v2");
    return DOLO_OK;
}
