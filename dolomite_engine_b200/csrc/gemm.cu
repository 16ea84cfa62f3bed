// bf16 GEMM for sm_100a:  TMA (cp.async.bulk.tensor) -> 128B-swizzled smem ring -> tcgen05.mma (UMMA 128x256x16,
// fp32 accumulators in TMEM, double buffered) -> epilogue warps (tcgen05.ld, alpha/bias/beta*C, bf16|fp32 store).
//
// Persistent, warp-specialised:  warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2..5 = epilogue.
// Both operands may be K-major (row-major [rows, K]) or MN-major (stored [K, rows]); the latter is what dgrad / wgrad
// need, so no transposes are ever materialised:
//     fwd   Y[T,N]  = X[T,K]  . W[N,K]^T          A K-major,  B K-major
//     dgrad dX[T,K] = dY[T,N] . W[N,K]            A K-major,  B MN-major (stored [N(contraction), K(out)])
//     wgrad dW[N,K] = dY[T,N]^T . X[T,K]          A MN-major, B MN-major (contraction over T)
#include <string.h>

#include "common.cuh"
#include "../../include/dolomite_b200.h"

using namespace dolo;

namespace {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle span
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int EPI_SLAB_BYTES = BM * 64 * 2;  // 128 rows x 64 bf16 columns, SW128
constexpr int EPI_BUFS = 2;
constexpr int GEMM_THREADS = 224;  // warp 0 TMA producer, 1 MMA issuer, 2..5 epilogue, 6 tile scheduler (dynamic mode)
constexpr int CLC_DEPTH = 8;       // tile-id responses in flight between the scheduler warp and the slowest role
constexpr int TMEM_COLS = 512;  // 2 accumulator stages x 256 fp32 columns

constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + EPI_BUFS * EPI_SLAB_BYTES + 512 /*barriers*/;
static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");

constexpr int MAXP = 4;  // problems per launch (the four weight gradients of a transformer block share one launch)

// Epilogue variants.  The fp32 TMA modes exist for the weight gradients: every epilogue warp stages its 32 rows x 32 fp32
// columns in a private shared-memory slab and hands it to the TMA engine as a tile store (first gradient of a window:
// overwrite) or a tile REDUCE-ADD performed by the L2 (accumulation) -- the first version read-modify-wrote the fp32
// accumulator with per-thread float4 accesses, one 128-byte line per lane.
enum EpiMode { EPI_DIRECT = 0, EPI_BF16_TMA = 1, EPI_F32_TMA_STORE = 2, EPI_F32_TMA_ADD = 3 };

struct GemmMaps {
    CUtensorMap a[MAXP], b[MAXP], d[MAXP];
};

struct Problem {
    void* D;
    const void* C;
    const __nv_bfloat16* bias;
    int64_t ldd, ldc;
    int M, N;
    int num_m, num_n, num_kb;
    int group_m;     // m-blocks per rasterisation panel
    int tile_start;  // first launch-wide tile index of this problem
    int epi;         // EpiMode
    float alpha, beta;
    uint64_t hint_a, hint_b;  // L2 eviction priority of the operand loads (TMA_HINT_*)
};

struct GemmParams {
    Problem pr[MAXP];
    int n_prob;
    int num_tiles;  // over all problems (grouped modes: single problem, includes the group factor)
    int dynamic;    // 1 = the grid has one cluster per tile and running clusters steal pending ones (cluster launch control)
    int d_is_f32;
    // grouped modes (MoE experts; moe_dolomite/moe/scatter.py:38-49 parallel_linear):
    //   1 = M-grouped: every 128-row tile of A/D belongs to one group (m_tile_group[m_blk], -1 = unused tile); B's outer
    //       TMA coordinate is offset by group * b_group_rows (fwd / dgrad of the expert linears)
    //   2 = K-grouped: tile index also enumerates the group; the contraction runs over rows
    //       [group_k_offsets[g], group_k_offsets[g+1]) and D/C are offset by g * d_group_stride (expert wgrad)
    int grouped;
    const int* m_tile_group;
    int m_tile_shift;        // CTA-pair mode: a 256-row super tile covers entries 2m, 2m+1 of the 128-row tile table
    const int* a_row_index;  // gather-on-load: source row of A for every (grouped) row of the problem, or NULL
    int b_group_rows;
    const int* group_k_offsets;
    int num_groups;
    int64_t d_group_stride;
};

struct TileInfo {
    int q;  // problem index
    int m_blk, n_blk, grp, kb0, kb1;
    bool valid;
};

// Tile rasterisation: sweep all n-blocks for a panel of `gm` m-blocks, so that the A panel (gm x 128 rows x K) stays
// L2-resident while B streams; gm is chosen on the host so that the panel is ~24 MB (ncu on the first version with a
// fixed panel of 8: B re-read 8x from DRAM, 870 MB of reads for 147 MB of operands).
__device__ __forceinline__ void tile_coords(int t, int num_m, int num_n, int gm, int& m_blk, int& n_blk) {
    const int per_group = gm * num_n;
    const int group = t / per_group;
    const int first_m = group * gm;
    const int gsize = min(num_m - first_m, gm);
    const int r = t - group * per_group;
    m_blk = first_m + r % gsize;
    n_blk = r / gsize;
}

__device__ __forceinline__ TileInfo tile_info(int t, const GemmParams& p) {
    TileInfo ti;
    ti.q = 0;
#pragma unroll
    for (int i = 1; i < MAXP; ++i)
        if (i < p.n_prob && t >= p.pr[i].tile_start) ti.q = i;
    const Problem& pr = p.pr[ti.q];
    t -= pr.tile_start;
    ti.grp = 0;
    ti.kb0 = 0;
    ti.kb1 = pr.num_kb;
    ti.valid = true;
    if (p.grouped == 3) {
        // split-K: tile index also enumerates the K split; partial products are reduced with fp32 atomics
        const int per = pr.num_m * pr.num_n;
        const int split = t / per;
        tile_coords(t - split * per, pr.num_m, pr.num_n, pr.group_m, ti.m_blk, ti.n_blk);
        const int kb_per = (pr.num_kb + p.num_groups - 1) / p.num_groups;
        ti.kb0 = split * kb_per;
        ti.kb1 = min(pr.num_kb, ti.kb0 + kb_per);
        ti.valid = ti.kb1 > ti.kb0;
    } else if (p.grouped == 2) {
        const int per = pr.num_m * pr.num_n;
        ti.grp = t / per;
        tile_coords(t - ti.grp * per, pr.num_m, pr.num_n, pr.group_m, ti.m_blk, ti.n_blk);
        ti.kb0 = p.group_k_offsets[ti.grp] / BK;
        ti.kb1 = p.group_k_offsets[ti.grp + 1] / BK;
        ti.valid = ti.kb1 > ti.kb0;
    } else {
        tile_coords(t, pr.num_m, pr.num_n, pr.group_m, ti.m_blk, ti.n_blk);
        if (p.grouped == 1) {
            ti.grp = p.m_tile_group[ti.m_blk << p.m_tile_shift];
            ti.valid = ti.grp >= 0;
        }
    }
    return ti;
}

// Where a role (producer / MMA issuer / epilogue warp) gets its next tile from.  Static mode: worker w takes tiles w,
// w + W, ...  Dynamic mode: the scheduler warp of the leader CTA keeps up to CLC_DEPTH cluster-launch-control requests
// ahead; every role of every CTA of the cluster consumes every response (the last one says "grid exhausted") and releases
// the slot on the LEADER's empty barrier.
template <bool CTA2>
struct TileFeed {
    const uint4* resp;
    uint64_t* full;
    uint64_t* empty;
    int slot;
    uint32_t phase;
    int step;  // static mode: number of workers
    bool dynamic;
    __device__ __forceinline__ TileFeed(const GemmParams& p, const uint4* r, uint64_t* f, uint64_t* e, int num_workers)
        : resp(r), full(f), empty(e), slot(0), phase(0), step(num_workers), dynamic(p.dynamic != 0) {}
    // `release` = this thread reports the slot as consumed for its role (one thread per role instance)
    __device__ __forceinline__ bool next(int& t, int num_tiles, bool release, bool whole_warp, int tag) {
        if (!dynamic) {
            t += step;
            return t < num_tiles;
        }
        mbar_wait(&full[slot], phase, tag);
        int x;
        const bool ok = clc_query(&resp[slot], x);
        fence_proxy_async_smem();  // the slot's next write comes from the async proxy
        if (whole_warp) __syncwarp();  // every lane has read the response before lane 0 hands the slot back
        if (release) {
            if constexpr (CTA2) mbar_arrive_remote(&empty[slot], 0);
            else mbar_arrive(&empty[slot]);
        }
        if (++slot == CLC_DEPTH) { slot = 0; phase ^= 1; }
        t = CTA2 ? (x >> 1) : x;
        return ok;
    }
};

// CTA2 = CTA-pair mode: a cluster of two CTAs computes a 256 x 256 tile with tcgen05.mma.cta_group::2 (M = 256).  Each
// CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256 N rows), so a pipeline stage is 32 KB instead
// of 48 KB (6 stages instead of 4) and B is fetched from L2 once per pair.  The MMA is issued by the leader CTA
// (cluster rank 0); accumulator rows 0..127 land in the leader's TMEM, rows 128..255 in the peer's, and each CTA runs its
// own epilogue.  In CTA2 mode p.num_m counts 256-row super tiles and grouped modes are not used.
template <bool A_MN, bool B_MN, bool CTA2>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_bf16_kernel(const __grid_constant__ GemmMaps maps, const __grid_constant__ GemmParams p) {
    constexpr int NSTAGE = CTA2 ? 6 : STAGES;
    constexpr int B_BYTES = CTA2 ? B_STAGE_BYTES / 2 : B_STAGE_BYTES;
    constexpr int STG_BYTES = A_STAGE_BYTES + B_BYTES;
    static_assert(NSTAGE * STG_BYTES == STAGES * STAGE_BYTES, "both modes use the same 192 KB operand ring");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_align_1024(smem_raw);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + NSTAGE * A_STAGE_BYTES;
    uint8_t* smem_epi = smem + NSTAGE * STG_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + EPI_BUFS * EPI_SLAB_BYTES);
    uint64_t* full_bar = bars;                   // [NSTAGE]
    uint64_t* empty_bar = bars + NSTAGE;         // [NSTAGE]
    uint64_t* tmem_full = bars + 2 * NSTAGE;     // [2]
    uint64_t* tmem_empty = bars + 2 * NSTAGE + 2;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 4);
    uint4* clc_resp = reinterpret_cast<uint4*>(bars + 24);  // [CLC_DEPTH] 16-byte responses (dynamic mode)
    uint64_t* clc_full = bars + 24 + 2 * CLC_DEPTH;         // [CLC_DEPTH]
    uint64_t* clc_empty = clc_full + CLC_DEPTH;             // [CLC_DEPTH] (the leader CTA's are the ones in use)
    static_assert(2 * NSTAGE + 5 <= 24 && (24 + 4 * CLC_DEPTH) * 8 <= 512, "barrier block layout");

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.num_tiles;
    const int cta_rank = CTA2 ? int(blockIdx.x & 1) : 0;          // == %cluster_ctarank for cluster dims (2,1,1)
    const int worker = CTA2 ? int(blockIdx.x >> 1) : int(blockIdx.x);  // persistent worker (CTA or CTA pair) index
                                                                       // == first tile in dynamic mode
    const int num_workers = CTA2 ? int(gridDim.x >> 1) : int(gridDim.x);

    if (warp == 0 && lane == 0) {
        for (int q = 0; q < p.n_prob; ++q) {
            tma_prefetch_desc(&maps.a[q]);
            tma_prefetch_desc(&maps.b[q]);
            if (p.pr[q].epi != EPI_DIRECT) tma_prefetch_desc(&maps.d[q]);
        }
        for (int i = 0; i < NSTAGE; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], CTA2 ? 8 : 4);  // one arrive per epilogue warp (of both CTAs in pair mode)
        }
        for (int i = 0; i < CLC_DEPTH; ++i) {
            mbar_init(&clc_full[i], 1);                // the scheduler's arrive.expect_tx (+ 16 response bytes)
            mbar_init(&clc_empty[i], CTA2 ? 11 : 6);   // producer(s) + MMA issuer + epilogue warps
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        if (CTA2) tmem_alloc_2cta<TMEM_COLS>(tmem_slot);
        else tmem_alloc<TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    if (CTA2) cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / multicast commit
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (!A_MN && p.a_row_index != nullptr) {
            // Gather-on-load (grouped expert GEMM reading the UNGROUPED activations): the whole warp takes part, lane l
            // owns rows 4l .. 4l+3 of the 128-row A tile and issues one gather4 per pipeline stage for them; lane 0 also
            // arms the barrier and loads B.
            int stage = 0;
            uint32_t phase = 0;
            for (int t = worker; t < num_tiles; t += num_workers) {
                const TileInfo ti = tile_info(t, p);
                if (!ti.valid) continue;
                const CUtensorMap* tmap_a = &maps.a[ti.q];
                const CUtensorMap* tmap_b = &maps.b[ti.q];
                const int m_blk = CTA2 ? ti.m_blk * 2 + cta_rank : ti.m_blk, n_blk = ti.n_blk;
                const int b_outer = (p.grouped == 1) ? ti.grp * p.b_group_rows : 0;
                const int4 src = *reinterpret_cast<const int4*>(p.a_row_index + int64_t(m_blk) * BM + lane * 4);
                for (int kb = ti.kb0; kb < ti.kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1, 1);
                    uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
                    uint8_t* sb = smem_b + stage * B_BYTES;
                    if (lane == 0) {
                        if constexpr (CTA2) {
                            if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * STG_BYTES);
                            const int n_row = n_blk * BN + cta_rank * (BN / 2);
                            if (!B_MN) {
                                tma_load_2d_2cta(sb, tmap_b, &full_bar[stage], kb * BK, b_outer + n_row);
                            } else {
#pragma unroll
                                for (int i = 0; i < BN / 128; ++i)
                                    tma_load_2d_2cta(sb + i * (BK * 128), tmap_b, &full_bar[stage], n_row + i * 64,
                                                     b_outer + kb * BK);
                            }
                        } else {
                            mbar_expect_tx(&full_bar[stage], STG_BYTES);
                            if (!B_MN) {
                                tma_load_2d(sb, tmap_b, &full_bar[stage], kb * BK, b_outer + n_blk * BN);
                            } else {
#pragma unroll
                                for (int i = 0; i < BN / 64; ++i)
                                    tma_load_2d(sb + i * (BK * 128), tmap_b, &full_bar[stage], n_blk * BN + i * 64,
                                                b_outer + kb * BK);
                            }
                        }
                    }
                    __syncwarp();
                    if constexpr (CTA2)
                        tma_gather4_2d_2cta(sa + lane * 512, tmap_a, &full_bar[stage], kb * BK, src.x, src.y, src.z, src.w);
                    else
                        tma_gather4_2d(sa + lane * 512, tmap_a, &full_bar[stage], kb * BK, src.x, src.y, src.z, src.w);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        } else if (elect_one()) {  // uniform single-thread region: ptxas keeps descriptors in uniform registers
            int stage = 0;
            uint32_t phase = 0;
            TileFeed<CTA2> feed(p, clc_resp, clc_full, clc_empty, num_workers);
            int t = worker, t_next = 0;
            bool have_next = false;
            for (bool have = t < num_tiles; have; t = t_next, have = have_next) {
                t_next = t;  // the next tile id is taken BEFORE this tile's work: its latency hides behind the pipeline waits
                have_next = feed.next(t_next, num_tiles, true, false, 11);
                const TileInfo ti = tile_info(t, p);
                if (!ti.valid) continue;
                const CUtensorMap* tmap_a = &maps.a[ti.q];
                const CUtensorMap* tmap_b = &maps.b[ti.q];
                const uint64_t ha = p.pr[ti.q].hint_a, hb = p.pr[ti.q].hint_b;
                const int m_blk = CTA2 ? ti.m_blk * 2 + cta_rank : ti.m_blk, n_blk = ti.n_blk;
                const int b_outer = (p.grouped == 1) ? ti.grp * p.b_group_rows : 0;
                for (int kb = ti.kb0; kb < ti.kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1, 1);
                    uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
                    uint8_t* sb = smem_b + stage * B_BYTES;
                    if constexpr (CTA2) {
                        // the leader's barrier collects the bytes of both CTAs (its own arrive carries the expectation)
                        if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * STG_BYTES);
                        const int n_row = n_blk * BN + cta_rank * (BN / 2);  // this CTA's half of the B tile
                        if (!A_MN) {
                            tma_load_2d_2cta_hint(sa, tmap_a, &full_bar[stage], kb * BK, m_blk * BM, ha);
                        } else {
#pragma unroll
                            for (int i = 0; i < BM / 64; ++i)
                                tma_load_2d_2cta_hint(sa + i * (BK * 128), tmap_a, &full_bar[stage], m_blk * BM + i * 64, kb * BK, ha);
                        }
                        if (!B_MN) {
                            tma_load_2d_2cta_hint(sb, tmap_b, &full_bar[stage], kb * BK, b_outer + n_row, hb);
                        } else {
#pragma unroll
                            for (int i = 0; i < BN / 128; ++i)
                                tma_load_2d_2cta_hint(sb + i * (BK * 128), tmap_b, &full_bar[stage], n_row + i * 64,
                                                 b_outer + kb * BK, hb);
                        }
                    } else {
                        mbar_expect_tx(&full_bar[stage], STG_BYTES);
                        if (!A_MN) {
                            tma_load_2d_hint(sa, tmap_a, &full_bar[stage], kb * BK, m_blk * BM, ha);
                        } else {
#pragma unroll
                            for (int i = 0; i < BM / 64; ++i)
                                tma_load_2d_hint(sa + i * (BK * 128), tmap_a, &full_bar[stage], m_blk * BM + i * 64, kb * BK, ha);
                        }
                        if (!B_MN) {
                            tma_load_2d_hint(sb, tmap_b, &full_bar[stage], kb * BK, b_outer + n_blk * BN, hb);
                        } else {
#pragma unroll
                            for (int i = 0; i < BN / 64; ++i)
                                tma_load_2d_hint(sb + i * (BK * 128), tmap_b, &full_bar[stage], n_blk * BN + i * 64,
                                            b_outer + kb * BK, hb);
                        }
                    }
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (cta_rank == 0 && elect_one()) {  // pair mode: only the leader CTA issues (for both SMs)
            constexpr uint32_t idesc = umma_idesc_bf16(CTA2 ? 2 * BM : BM, BN, A_MN, B_MN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            TileFeed<CTA2> feed(p, clc_resp, clc_full, clc_empty, num_workers);
            int t = worker, t_next = 0;
            bool have_next = false;
            for (bool have = t < num_tiles; have; t = t_next, have = have_next) {
                t_next = t;  // the next tile id is taken BEFORE this tile's work: its latency hides behind the pipeline waits
                have_next = feed.next(t_next, num_tiles, true, false, 12);
                const TileInfo ti = tile_info(t, p);
                if (!ti.valid) continue;
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 2);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);
                for (int kb = ti.kb0; kb < ti.kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase, 3);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem_a + stage * A_STAGE_BYTES);
                    const uint32_t sb = smem_u32(smem_b + stage * B_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // K-major: +32 B per K=16 step inside the 128 B swizzle span; SBO = 8 rows x 128 B.
                        // MN-major: +16 K-rows x 128 B per step; LBO = next 64-wide MN chunk (BK rows x 128 B), SBO = 8 K-rows.
                        // (pair mode: the same descriptors address each CTA's own A rows / B half at equal smem offsets)
                        const uint64_t adesc = A_MN ? umma_smem_desc(sa + k * 2048, BK * 128, 1024, 2)
                                                    : umma_smem_desc(sa + k * 32, 16, 1024, 2);
                        const uint64_t bdesc = B_MN ? umma_smem_desc(sb + k * 2048, BK * 128, 1024, 2)
                                                    : umma_smem_desc(sb + k * 32, 16, 1024, 2);
                        if constexpr (CTA2) umma_ss_2cta(d_tmem, adesc, bdesc, idesc, (kb != ti.kb0 || k != 0) ? 1u : 0u);
                        else umma_ss(d_tmem, adesc, bdesc, idesc, (kb != ti.kb0 || k != 0) ? 1u : 0u);
                    }
                    // smem slot reusable (in both CTAs) once these MMAs retire
                    if constexpr (CTA2) umma_commit_2cta(&empty_bar[stage]);
                    else umma_commit(&empty_bar[stage]);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
                // accumulator complete (each CTA's epilogue drains its own 128 rows)
                if constexpr (CTA2) umma_commit_2cta(&tmem_full[acc]);
                else umma_commit(&tmem_full[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp == 6) {
        // ================= tile scheduler (dynamic mode; leader CTA only) =================
        if (p.dynamic != 0 && cta_rank == 0 && elect_one()) {
            int slot = 0;
            uint32_t phase = 0;
            for (;;) {
                mbar_wait(&clc_empty[slot], phase ^ 1, 20);  // every role of both CTAs has read the previous response
                mbar_expect_tx(&clc_full[slot], 16);
                if constexpr (CTA2) {
                    mbar_expect_tx_remote(&clc_full[slot], 16, 1);
                    clc_try_cancel_multicast(&clc_resp[slot], &clc_full[slot]);
                } else {
                    clc_try_cancel(&clc_resp[slot], &clc_full[slot]);
                }
                mbar_wait(&clc_full[slot], phase, 21);
                int x;
                const bool ok = clc_query(&clc_resp[slot], x);
                fence_proxy_async_smem();
                if (++slot == CLC_DEPTH) { slot = 0; phase ^= 1; }
                if (!ok) break;  // no request may follow a failed one
            }
        }
    } else {
        // ================= epilogue (4 warps, TMEM sub-partition = warp % 4) =================
        const int sub = warp & 3;
        const int et = sub * 32 + lane;  // row within the tile == TMEM lane
        int acc = 0;
        uint32_t acc_phase = 0;
        int epi_buf = 0;
        bool cta_stores = false, warp_stores = false;  // which kind of bulk group this thread may still have in flight
        uint8_t* wslab = smem_epi + sub * (EPI_BUFS * EPI_SLAB_BYTES / 4);  // this warp's two private 4 KB slabs (fp32 modes)
        int wbuf = 0;
        TileFeed<CTA2> feed(p, clc_resp, clc_full, clc_empty, num_workers);
        int t = worker, t_next = 0;
        bool have_next = false;
        for (bool have = t < num_tiles; have; t = t_next, have = have_next) {
            t_next = t;  // the next tile id is taken BEFORE this tile's work: its latency hides behind the pipeline waits
            have_next = feed.next(t_next, num_tiles, lane == 0, true, 13);
            const TileInfo ti = tile_info(t, p);
            if (!ti.valid) {
                // K-grouped launch (expert weight gradients) and this expert received NO rows: its product is zero.  A launch
                // that OVERWRITES (beta = 0, no C) must still write the tile -- the caller did not clear the buffer -- so the
                // epilogue warps store zeros themselves; no MMA ran and no accumulator stage is involved.
                const Problem& pz = p.pr[ti.q];
                if (p.grouped == 2 && p.d_is_f32 && pz.C == nullptr) {
                    const int64_t zrow = int64_t(CTA2 ? ti.m_blk * 2 + cta_rank : ti.m_blk) * BM + et;
                    if (zrow < pz.M) {
                        float* zr = static_cast<float*>(pz.D) + int64_t(ti.grp) * p.d_group_stride + zrow * pz.ldd;
                        const int c1 = min(pz.N, (ti.n_blk + 1) * BN);
                        for (int c = ti.n_blk * BN; c < c1; c += 4)  // N % 8 == 0: whole float4 in range
                            *reinterpret_cast<float4*>(zr + c) = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                continue;
            }
            const Problem& pr = p.pr[ti.q];
            const CUtensorMap* tmap_d = &maps.d[ti.q];
            const int m_blk = CTA2 ? ti.m_blk * 2 + cta_rank : ti.m_blk, n_blk = ti.n_blk;
            const int64_t d_off = (p.grouped == 2) ? int64_t(ti.grp) * p.d_group_stride : 0;
            mbar_wait(&tmem_full[acc], acc_phase, 4);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (uint32_t(sub * 32) << 16) + uint32_t(acc * BN);
            const int64_t row = int64_t(m_blk) * BM + et;
            const bool row_ok = row < pr.M;
            const int col0 = n_blk * BN;
            const float alpha = pr.alpha, beta = pr.beta;
            const __nv_bfloat16* bias = pr.bias;
            const int N = pr.N;

            if (pr.epi == EPI_BF16_TMA) {
                cta_stores = true;
                // 4 slabs of 64 columns: regs -> swizzled smem -> TMA store
#pragma unroll 1
                for (int slab = 0; slab < BN / 64; ++slab) {
                    if (col0 + slab * 64 >= N) break;
                    uint8_t* buf = smem_epi + epi_buf * EPI_SLAB_BYTES;
                    // the buffer must have been fully read by the TMA store issued two slabs ago
                    if (et == 0) tma_store_wait_read<EPI_BUFS - 1>();
                    named_bar_sync(1, 128);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t r[32];
                        tmem_ld32(t_addr + slab * 64 + h * 32, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 4; ++c) {  // 4 x 16-byte chunks (8 columns each)
                            float f[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(r[c * 8 + j]);
                            if (bias != nullptr) {
                                const int cb = col0 + slab * 64 + h * 32 + c * 8;
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    if (cb + j < N) f[j] += __bfloat162float(bias[cb + j]);
                            }
#pragma unroll
                            for (int j = 0; j < 8; ++j) f[j] *= alpha;
                            uint4 v;
                            v.x = pack_bf16(f[0], f[1]);
                            v.y = pack_bf16(f[2], f[3]);
                            v.z = pack_bf16(f[4], f[5]);
                            v.w = pack_bf16(f[6], f[7]);
                            const int chunk = h * 4 + c;  // logical 16-byte chunk inside the 128-byte row
                            *reinterpret_cast<uint4*>(buf + et * 128 + ((chunk ^ (et & 7)) << 4)) = v;
                        }
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(1, 128);
                    if (et == 0) {
                        tma_store_2d(tmap_d, buf, col0 + slab * 64, m_blk * BM);
                        tma_store_commit();
                    }
                    epi_buf ^= 1;
                }
            } else if (pr.epi == EPI_F32_TMA_STORE || pr.epi == EPI_F32_TMA_ADD) {
                // fp32 D (weight gradients): 8 slabs of 32 columns.  Each warp owns its 32 rows: TMEM -> registers -> a
                // private SW128 slab (32 rows x 128 B) -> one TMA tile store / reduce-add per slab.  No CTA-wide barrier.
                warp_stores = true;
                const bool add = pr.epi == EPI_F32_TMA_ADD;
                const int row0 = m_blk * BM + sub * 32;
#pragma unroll 1
                for (int slab = 0; slab < BN / 32; ++slab) {
                    const int cb = col0 + slab * 32;
                    if (cb >= N) break;
                    uint8_t* buf = wslab + wbuf * 4096;
                    if (lane == 0) tma_store_wait_read<1>();  // the operation issued from this slab two slabs ago has read it
                    __syncwarp();
                    uint32_t r[32];
                    tmem_ld32(t_addr + slab * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        uint4 v;
                        v.x = __float_as_uint(__uint_as_float(r[c * 4 + 0]) * alpha);
                        v.y = __float_as_uint(__uint_as_float(r[c * 4 + 1]) * alpha);
                        v.z = __float_as_uint(__uint_as_float(r[c * 4 + 2]) * alpha);
                        v.w = __float_as_uint(__uint_as_float(r[c * 4 + 3]) * alpha);
                        *reinterpret_cast<uint4*>(buf + lane * 128 + ((c ^ (lane & 7)) << 4)) = v;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        if (row0 < pr.M) {  // rows / columns beyond the tensor are clipped by the TMA unit
                            if (add) tma_reduce_add_3d(tmap_d, buf, cb, row0, ti.grp);
                            else tma_store_3d(tmap_d, buf, cb, row0, ti.grp);
                        }
                        tma_store_commit();
                    }
                    wbuf ^= 1;
                }
            } else {
#pragma unroll 1
                for (int ch = 0; ch < BN / 32; ++ch) {
                    const int cb = col0 + ch * 32;
                    if (cb >= N) break;
                    uint32_t r[32];
                    tmem_ld32(t_addr + ch * 32, r);
                    tmem_ld_wait();
                    if (!row_ok) continue;
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(r[j]);
                    if (bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (cb + j < N) f[j] += __bfloat162float(bias[cb + j]);
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] *= alpha;
                    if (p.d_is_f32) {
                        float* drow = static_cast<float*>(pr.D) + d_off + row * pr.ldd + cb;
                        const float* crow = pr.C ? static_cast<const float*>(pr.C) + d_off + row * pr.ldc + cb : nullptr;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            if (cb + q * 4 < N) {  // N % 8 == 0 -> whole float4 in range
                                float4 o = make_float4(f[q * 4], f[q * 4 + 1], f[q * 4 + 2], f[q * 4 + 3]);
                                if (p.grouped == 3) {  // split-K partial: D += o
                                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + q * 4),
                                                 "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w)
                                                 : "memory");
                                    continue;
                                }
                                if (crow) {
                                    const float4 c4 = *reinterpret_cast<const float4*>(crow + q * 4);
                                    o.x += beta * c4.x; o.y += beta * c4.y; o.z += beta * c4.z; o.w += beta * c4.w;
                                }
                                *reinterpret_cast<float4*>(drow + q * 4) = o;
                            }
                        }
                    } else {
                        __nv_bfloat16* drow = static_cast<__nv_bfloat16*>(pr.D) + d_off + row * pr.ldd + cb;
                        const __nv_bfloat16* crow =
                            pr.C ? static_cast<const __nv_bfloat16*>(pr.C) + d_off + row * pr.ldc + cb : nullptr;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (cb + q * 8 < N) {
                                if (crow) {
                                    const uint4 cv = *reinterpret_cast<const uint4*>(crow + q * 8);
                                    f[q * 8 + 0] += beta * bf16_lo(cv.x); f[q * 8 + 1] += beta * bf16_hi(cv.x);
                                    f[q * 8 + 2] += beta * bf16_lo(cv.y); f[q * 8 + 3] += beta * bf16_hi(cv.y);
                                    f[q * 8 + 4] += beta * bf16_lo(cv.z); f[q * 8 + 5] += beta * bf16_hi(cv.z);
                                    f[q * 8 + 6] += beta * bf16_lo(cv.w); f[q * 8 + 7] += beta * bf16_hi(cv.w);
                                }
                                uint4 v;
                                v.x = pack_bf16(f[q * 8 + 0], f[q * 8 + 1]);
                                v.y = pack_bf16(f[q * 8 + 2], f[q * 8 + 3]);
                                v.z = pack_bf16(f[q * 8 + 4], f[q * 8 + 5]);
                                v.w = pack_bf16(f[q * 8 + 6], f[q * 8 + 7]);
                                *reinterpret_cast<uint4*>(drow + q * 8) = v;
                            }
                        }
                    }
                }
            }
            // release the accumulator stage back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CTA2) mbar_arrive_remote(&tmem_empty[acc], 0);  // the leader's MMA thread waits for both CTAs
                else mbar_arrive(&tmem_empty[acc]);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        // shared memory must outlive every bulk operation that still reads it
        if ((cta_stores && et == 0) || (warp_stores && lane == 0)) tma_store_wait_all<0>();
    }

    tc_fence_before();
    if constexpr (CTA2) {
        cluster_sync_all();  // the peer's smem / TMEM / barriers stay valid until both CTAs are done
        if (warp == 1) {
            tc_fence_after();
            tmem_dealloc_2cta<TMEM_COLS>(tmem_base);
        }
    } else {
        __syncthreads();
        if (warp == 1) {
            tc_fence_after();
            tmem_dealloc<TMEM_COLS>(tmem_base);
        }
    }
}

template <bool A_MN, bool B_MN>
int launch_gemm(const GemmMaps& maps, const GemmParams& p_in, cudaStream_t st, bool cta_pair) {
    // Dynamic mode: one cluster per tile, running clusters cancel + take over pending ones (TileFeed).  The grid then
    // uses whatever SMs are free when it starts -- and the ones that get free while it runs -- so gemm_sm_margin is moot.
    GemmParams p = p_in;
    p.dynamic = (dolo_option_gemm_dynamic() != 0 && p.a_row_index == nullptr) ? 1 : 0;
    const int tiles = p.num_tiles;
    if (cta_pair) {
        auto kern = gemm_bf16_kernel<A_MN, B_MN, true>;
        static bool attr_set2 = false;  // per instantiation
        if (!attr_set2) {
            DOLO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            attr_set2 = true;
        }
        const int pairs = (dolo_num_sms() - dolo_option_gemm_sm_margin()) / 2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(unsigned(2 * ((p.dynamic || tiles < pairs) ? tiles : pairs)));
        cfg.blockDim = dim3(GEMM_THREADS);
        cfg.dynamicSmemBytes = SMEM_BYTES;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        DOLO_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, maps, p));
        return DOLO_OK;
    }
    auto kern = gemm_bf16_kernel<A_MN, B_MN, false>;
    static bool attr_set = false;  // per instantiation
    if (!attr_set) {
        DOLO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    const int sms = dolo_num_sms() - dolo_option_gemm_sm_margin();
    const int grid = (p.dynamic || tiles < sms) ? tiles : sms;
    kern<<<grid, GEMM_THREADS, SMEM_BYTES, st>>>(maps, p);
    DOLO_LAUNCH_OK("gemm_bf16");
    return DOLO_OK;
}

}  // namespace

struct GroupArgs {
    int mode = 0;
    const int* a_row_index = nullptr;  // mode 1 only: gather the rows of A on load (A is the ungrouped matrix of a_rows rows)
    int64_t a_rows = 0;
    const int* m_tile_group = nullptr;
    int64_t b_group_rows = 0;
    const int* group_k_offsets = nullptr;
    int num_groups = 1;
    int64_t d_group_stride = 0;
    int64_t b_total_outer = 0;  // rows of B's outer TMA dimension over all groups
};

// one problem of a launch
struct GemmProblemArgs {
    const void* A; int64_t lda;
    const void* B; int64_t ldb;
    void* D; int64_t ldd;
    const void* C; int64_t ldc;
    const void* bias;
    float alpha, beta;
    int64_t M, N, K;
};

// Fills maps.{a,b,d}[q] and p.pr[q] for one problem.  All problems of a launch share the operand layouts, the output type,
// the epilogue kind and the CTA-pair decision.
static int setup_problem(GemmMaps& maps, GemmParams& p, int q, const GemmProblemArgs& g, int a_mn_major, int b_mn_major,
                         int d_is_f32, int epi, bool cta_pair, const GroupArgs& ga) {
    const int64_t M = g.M, N = g.N, K = g.K;
    DOLO_REQUIRE(K > 0, "gemm: K must be > 0");
    DOLO_REQUIRE(K % 8 == 0 && N % 8 == 0, "gemm: K=%lld and N=%lld must be multiples of 8", (long long)K, (long long)N);
    DOLO_REQUIRE(g.lda % 8 == 0 && g.ldb % 8 == 0 && g.ldd % (d_is_f32 ? 4 : 8) == 0,
                 "gemm: leading dimensions must keep 16-byte alignment");
    DOLO_REQUIRE(!a_mn_major || M % 8 == 0, "gemm: MN-major A requires M %% 8 == 0");
    DOLO_REQUIRE(g.C == nullptr || g.ldc % (d_is_f32 ? 4 : 8) == 0, "gemm: ldc alignment");
    DOLO_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "gemm: dimension too large");
    // K-major: dims {K, rows}, box {64, tile rows}.  MN-major: dims {rows, K}, box {64, 64}.
    uint64_t dims[3], strides[3];
    uint32_t box[3];
    if (ga.a_row_index != nullptr) {
        // gather4: box = one row of 64 columns; the instruction names four rows of the UNGROUPED matrix
        DOLO_REQUIRE(!a_mn_major, "gemm: gather-on-load needs a K-major A");
        dims[0] = uint64_t(K); dims[1] = uint64_t(ga.a_rows); strides[0] = 2; strides[1] = uint64_t(g.lda) * 2;
        box[0] = BK; box[1] = 1;
    } else if (!a_mn_major) {
        dims[0] = uint64_t(K); dims[1] = uint64_t(M); strides[0] = 2; strides[1] = uint64_t(g.lda) * 2;
        box[0] = BK; box[1] = BM;
    } else {
        dims[0] = uint64_t(M); dims[1] = uint64_t(K); strides[0] = 2; strides[1] = uint64_t(g.lda) * 2;
        box[0] = 64; box[1] = BK;
    }
    int rc = dolo_make_tmap(&maps.a[q], g.A, 2, 2, dims, strides, box, DOLO_SW_128);
    if (rc) return rc;
    if (!b_mn_major) {
        dims[0] = uint64_t(K); dims[1] = uint64_t(ga.mode == 1 ? ga.b_total_outer : N); strides[1] = uint64_t(g.ldb) * 2;
        box[0] = BK; box[1] = cta_pair ? BN / 2 : BN;  // pair mode: each CTA stages half of the B tile
    } else {
        dims[0] = uint64_t(N); dims[1] = uint64_t(ga.mode == 1 ? ga.b_total_outer : K); strides[1] = uint64_t(g.ldb) * 2;
        box[0] = 64; box[1] = BK;
    }
    rc = dolo_make_tmap(&maps.b[q], g.B, 2, 2, dims, strides, box, DOLO_SW_128);
    if (rc) return rc;
    if (epi == EPI_BF16_TMA) {
        dims[0] = uint64_t(N); dims[1] = uint64_t(M); strides[1] = uint64_t(g.ldd) * 2;
        box[0] = 64; box[1] = BM;
        rc = dolo_make_tmap(&maps.d[q], g.D, 2, 2, dims, strides, box, DOLO_SW_128);
        if (rc) return rc;
    } else if (epi == EPI_F32_TMA_STORE || epi == EPI_F32_TMA_ADD) {
        // fp32 [groups, M, N]: one box = 32 columns (128 B) x 32 rows, the slab of one epilogue warp
        const int64_t groups = ga.mode == 2 ? ga.num_groups : 1;
        dims[0] = uint64_t(N); dims[1] = uint64_t(M); dims[2] = uint64_t(groups);
        strides[1] = uint64_t(g.ldd) * 4;
        strides[2] = uint64_t(ga.mode == 2 ? ga.d_group_stride : M * g.ldd) * 4;
        box[0] = 32; box[1] = 32; box[2] = 1;
        rc = dolo_make_tmap(&maps.d[q], g.D, 4, 3, dims, strides, box, DOLO_SW_128);
        if (rc) return rc;
    } else {
        maps.d[q] = maps.a[q];  // unused
    }
    Problem& pr = p.pr[q];
    pr.D = g.D;
    pr.C = g.C;
    pr.bias = static_cast<const __nv_bfloat16*>(g.bias);
    pr.ldd = g.ldd;
    pr.ldc = g.ldc;
    pr.M = int(M);
    pr.N = int(N);
    pr.alpha = g.alpha;
    pr.beta = g.C ? g.beta : 0.f;
    pr.epi = epi;
    pr.hint_a = pr.hint_b = TMA_HINT_NORMAL;
    if (dolo_option_gemm_l2_hints() && ga.mode == 0 && K >= 4096 && (M + N) * K * 2 > (48ll << 20)) {
        // long contraction, operands larger than what the L2 keeps anyway: stream the bigger one, keep the smaller one
        const bool a_smaller = M <= N;
        pr.hint_a = a_smaller ? TMA_HINT_EVICT_LAST : TMA_HINT_EVICT_FIRST;
        pr.hint_b = a_smaller ? TMA_HINT_EVICT_FIRST : TMA_HINT_EVICT_LAST;
    }
    pr.num_m = cta_pair ? int((M + 2 * BM - 1) / (2 * BM)) : int((M + BM - 1) / BM);  // pair mode: 256-row super tiles
    pr.num_n = int((N + BN - 1) / BN);
    pr.num_kb = int((K + BK - 1) / BK);
    {
        // A panel of group_m x 128 rows x K bf16 should fit comfortably in L2 next to the streaming B tiles
        const int64_t panel_bytes = int64_t(cta_pair ? 2 * BM : BM) * K * 2;
        int64_t gm = (24ll << 20) / (panel_bytes > 0 ? panel_bytes : 1);
        if (gm < 4) gm = 4;
        if (gm > 64) gm = 64;
        pr.group_m = int(gm);
    }
    return DOLO_OK;
}

template <typename... Ts>
static int dispatch_layout(int a_mn_major, int b_mn_major, Ts&&... args) {
    if (!a_mn_major && !b_mn_major) return launch_gemm<false, false>(args...);
    if (!a_mn_major && b_mn_major) return launch_gemm<false, true>(args...);
    if (a_mn_major && !b_mn_major) return launch_gemm<true, false>(args...);
    return launch_gemm<true, true>(args...);
}

// fp32 outputs go through the TMA epilogue when they are a plain overwrite (no C) or an in-place accumulation
// (C == D, beta == 1) without bias: exactly the two forms a weight gradient takes
static int pick_epilogue(int d_is_f32, const void* C, const void* D, float beta, const void* bias, bool tma_store,
                         int group_mode) {
    if (!d_is_f32) return tma_store ? EPI_BF16_TMA : EPI_DIRECT;
    if (bias != nullptr || group_mode == 3 || !dolo_option_gemm_f32_tma_epilogue()) return EPI_DIRECT;
    if (C == nullptr) return EPI_F32_TMA_STORE;
    if (C == D && beta == 1.f) return EPI_F32_TMA_ADD;
    return EPI_DIRECT;
}

static int gemm_impl(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major, void* D,
                     int64_t ldd, int d_is_f32, const void* C, int64_t ldc, float alpha, float beta, const void* bias,
                     int64_t M, int64_t N, int64_t K, int flags, void* stream, const GroupArgs& ga) {
    DOLO_REQUIRE(M >= 0 && N >= 0 && K >= 0, "gemm: negative dimension");
    if (M == 0 || N == 0) return DOLO_OK;
    const bool tma_store = (flags & DOLO_GEMM_FLAG_TMA_STORE) != 0;
    DOLO_REQUIRE(!tma_store || (!d_is_f32 && C == nullptr), "gemm: TMA-store epilogue needs bf16 D and no C");
    // CTA-pair (cta_group::2) kernel: at least one full 256-row super tile; M-grouped problems need their expert segments
    // padded to 256 rows (both halves of a super tile then belong to the same expert); not for split-K
    const bool cta_pair = (ga.mode == 0 || (ga.mode == 1 && M % (2 * BM) == 0) || ga.mode == 2) && M >= 2 * BM &&
                          ((flags & DOLO_GEMM_FLAG_CTA_PAIR) != 0 || dolo_option_gemm_cta_pair() != 0) &&
                          (flags & DOLO_GEMM_FLAG_NO_CTA_PAIR) == 0;
    int epi = pick_epilogue(d_is_f32, C, D, beta, bias, tma_store, ga.mode);
    if ((flags & DOLO_GEMM_FLAG_DIRECT_EPILOGUE) != 0 && d_is_f32) epi = EPI_DIRECT;
    if ((flags & DOLO_GEMM_FLAG_F32_TMA_EPILOGUE) != 0 && d_is_f32 && bias == nullptr && ga.mode != 3)
        epi = C == nullptr ? EPI_F32_TMA_STORE : ((C == D && beta == 1.f) ? EPI_F32_TMA_ADD : EPI_DIRECT);
    GemmMaps maps;
    GemmParams p;
    memset(&p, 0, sizeof(p));
    GemmProblemArgs g{A, lda, B, ldb, D, ldd, C, ldc, bias, alpha, beta, M, N, K};
    int rc = setup_problem(maps, p, 0, g, a_mn_major, b_mn_major, d_is_f32, epi, cta_pair, ga);
    if (rc) return rc;
    p.n_prob = 1;
    p.pr[0].tile_start = 0;
    p.num_tiles = p.pr[0].num_m * p.pr[0].num_n * (ga.mode >= 2 ? ga.num_groups : 1);
    p.d_is_f32 = d_is_f32;
    p.grouped = ga.mode;
    p.m_tile_group = ga.m_tile_group;
    p.m_tile_shift = cta_pair ? 1 : 0;
    p.a_row_index = ga.a_row_index;
    p.b_group_rows = int(ga.b_group_rows);
    p.group_k_offsets = ga.group_k_offsets;
    p.num_groups = ga.num_groups;
    p.d_group_stride = ga.d_group_stride;
    return dispatch_layout(a_mn_major, b_mn_major, maps, p, static_cast<cudaStream_t>(stream), cta_pair);
}

// The weight gradients of one transformer block in ONE persistent launch (autograd of linear.py:5-25 for c_attn, attention
// c_proj, c_fc and mlp c_proj): dW_i[M_i, N_i] (+)= alpha_i * dY_i^T X_i with dY_i [K, M_i], X_i [K, N_i] row-major
// activations (both operands MN-major).  Launched one by one these GEMMs lose 10-30 % to wave quantisation (c_attn: 300
// pair tiles on 74 CTA pairs = 4.05 waves; attention c_proj: 1.35 waves); together they are 1600 tiles = 21.6 waves.
extern "C" int dolomite_b200_gemm_bf16_wgrad_multi(int n_problems, const void* const* dY, const int64_t* ld_dy,
                                                   const void* const* X, const int64_t* ld_x, float* const* dW,
                                                   const int64_t* ld_dw, const int64_t* M, const int64_t* N, int64_t K,
                                                   const float* alpha, const int* accumulate, void* stream) {
    DOLO_REQUIRE(n_problems >= 1 && n_problems <= MAXP, "wgrad_multi: between 1 and %d problems per launch", MAXP);
    GemmMaps maps;
    GemmParams p;
    memset(&p, 0, sizeof(p));
    bool pair = dolo_option_gemm_cta_pair() != 0;
    for (int q = 0; q < n_problems; ++q) pair = pair && M[q] >= 2 * BM;
    int tiles = 0;
    for (int q = 0; q < n_problems; ++q) {
        DOLO_REQUIRE(M[q] > 0 && N[q] > 0, "wgrad_multi: empty problem %d", q);
        GemmProblemArgs g{dY[q], ld_dy[q], X[q], ld_x[q], dW[q], ld_dw[q], accumulate[q] ? dW[q] : nullptr, ld_dw[q], nullptr,
                          alpha[q], 1.f, M[q], N[q], K};
        const int epi = !dolo_option_gemm_f32_tma_epilogue() ? EPI_DIRECT : (accumulate[q] ? EPI_F32_TMA_ADD : EPI_F32_TMA_STORE);
        int rc = setup_problem(maps, p, q, g, 1, 1, 1, epi, pair, GroupArgs());
        if (rc) return rc;
        p.pr[q].tile_start = tiles;
        tiles += p.pr[q].num_m * p.pr[q].num_n;
    }
    p.n_prob = n_problems;
    p.num_tiles = tiles;
    p.d_is_f32 = 1;
    p.grouped = 0;
    p.num_groups = 1;
    return launch_gemm<true, true>(maps, p, static_cast<cudaStream_t>(stream), pair);
}

extern "C" int dolomite_b200_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb,
                                       int b_mn_major, void* D, int64_t ldd, int d_is_f32, const void* C, int64_t ldc,
                                       float alpha, float beta, const void* bias, int64_t M, int64_t N, int64_t K,
                                       int flags, void* stream) {
    if (flags & DOLO_GEMM_FLAG_SPLITK_ACCUMULATE) {
        // D(fp32) += alpha * A B^T with the contraction split over several CTAs (fp32 vector atomics): removes the
        // wave-quantisation tail of weight-gradient GEMMs (few output tiles, long K) and the read of C
        DOLO_REQUIRE(d_is_f32 && bias == nullptr && (C == nullptr || (C == D && beta == 1.f)),
                     "gemm: split-K accumulate needs fp32 D, no bias and C == D with beta == 1");
        const int64_t tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
        const int64_t num_kb = (K + BK - 1) / BK;
        const int sms = dolo_num_sms();
        int best_s = 1;
        double best = 1e30;
        for (int s = 1; s <= 16; ++s) {
            if (s > 1 && num_kb / s < 8) break;
            const double waves = double((tiles * s + sms - 1) / sms) / double(s);  // in units of one full-K tile time
            if (waves < best - 1e-9) { best = waves; best_s = s; }
        }
        GroupArgs ga;
        ga.mode = 3;
        ga.num_groups = best_s;
        return gemm_impl(A, lda, a_mn_major, B, ldb, b_mn_major, D, ldd, 1, nullptr, 0, alpha, 0.f, nullptr, M, N, K, 0,
                         stream, ga);
    }
    return gemm_impl(A, lda, a_mn_major, B, ldb, b_mn_major, D, ldd, d_is_f32, C, ldc, alpha, beta, bias, M, N, K, flags,
                     stream, GroupArgs());
}

extern "C" int dolomite_b200_gemm_bf16_grouped_m(const void* A, int64_t lda, const void* B, int64_t ldb, int b_mn_major,
                                                 void* D, int64_t ldd, float alpha, int64_t M_max, int64_t N, int64_t K,
                                                 const int32_t* m_tile_group, int num_groups, int flags, void* stream) {
    DOLO_REQUIRE(M_max % BM == 0, "grouped gemm: M_max=%lld must be a multiple of %d (padded expert segments)",
                 (long long)M_max, BM);
    DOLO_REQUIRE(!b_mn_major || K % BK == 0, "grouped gemm: MN-major B needs K %% %d == 0", BK);
    DOLO_REQUIRE(m_tile_group != nullptr && num_groups > 0, "grouped gemm: missing group table");
    GroupArgs ga;
    ga.mode = 1;
    ga.m_tile_group = m_tile_group;
    ga.num_groups = num_groups;
    ga.b_group_rows = b_mn_major ? K : N;
    ga.b_total_outer = ga.b_group_rows * num_groups;
    return gemm_impl(A, lda, 0, B, ldb, b_mn_major, D, ldd, 0, nullptr, 0, alpha, 0.f, nullptr, M_max, N, K, flags, stream,
                     ga);
}

// Same, with the ScatterMoE gather fused into the operand load: A is the UNGROUPED activation matrix [a_rows, K] and
// a_row_index[r] names the source row of grouped row r (padding rows may name any valid row: their products are never read)
extern "C" int dolomite_b200_gemm_bf16_grouped_m_gather(const void* A, int64_t lda, int64_t a_rows,
                                                        const int32_t* a_row_index, const void* B, int64_t ldb, void* D,
                                                        int64_t ldd, float alpha, int64_t M_max, int64_t N, int64_t K,
                                                        const int32_t* m_tile_group, int num_groups, int flags,
                                                        void* stream) {
    DOLO_REQUIRE(M_max % BM == 0, "grouped gemm: M_max=%lld must be a multiple of %d (padded expert segments)",
                 (long long)M_max, BM);
    DOLO_REQUIRE(m_tile_group != nullptr && num_groups > 0 && a_row_index != nullptr && a_rows > 0,
                 "grouped gemm (gather): missing group table / row index");
    DOLO_REQUIRE((reinterpret_cast<uintptr_t>(a_row_index) & 15) == 0, "grouped gemm (gather): row index must be 16-byte aligned");
    GroupArgs ga;
    ga.mode = 1;
    ga.a_row_index = a_row_index;
    ga.a_rows = a_rows;
    ga.m_tile_group = m_tile_group;
    ga.num_groups = num_groups;
    ga.b_group_rows = N;
    ga.b_total_outer = N * num_groups;
    return gemm_impl(A, lda, 0, B, ldb, 0, D, ldd, 0, nullptr, 0, alpha, 0.f, nullptr, M_max, N, K, flags, stream, ga);
}

extern "C" int dolomite_b200_gemm_bf16_grouped_k(const void* A, int64_t lda, const void* B, int64_t ldb, float* D,
                                                 int64_t ldd, float alpha, float beta, int64_t M, int64_t N,
                                                 int64_t K_max, const int32_t* group_k_offsets, int num_groups,
                                                 void* stream) {
    DOLO_REQUIRE(group_k_offsets != nullptr && num_groups > 0, "grouped wgrad: missing offsets");
    GroupArgs ga;
    ga.mode = 2;
    ga.group_k_offsets = group_k_offsets;
    ga.num_groups = num_groups;
    ga.d_group_stride = M * ldd;
    // A, B are both MN-major views of [K_max, M] / [K_max, N] row-major activations; D[g] (+)= A_g^T B_g in fp32
    return gemm_impl(A, lda, 1, B, ldb, 1, D, ldd, 1, beta != 0.f ? D : nullptr, ldd, alpha, beta, nullptr, M, N, K_max, 0,
                     stream, ga);
}
