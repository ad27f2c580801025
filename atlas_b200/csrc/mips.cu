// Exact max-inner-product search with fused top-k candidate filtering for sm_100a.
//
// Replaces the reference's `torch.matmul(q.half(), E)` + `torch.topk` (src/index.py:113-120): the
// [nq, n] score matrix is never written.  Pipeline per block of <= 256 queries:
//
//   (1) sample scan   : the scan kernel over a strided sample of ~2*sqrt(k*n) bank rows, every
//                       score kept                                   -> candidate lists
//   (2) select (thr)  : per query, exact k-th best of the sample      -> a VALID lower bound on the
//                       final k-th best score (k real passages reach it)
//   (3) main scan     : ONE sweep of the whole bank (the HBM-bound part): tcgen05 GEMM tiles
//                       [<=256 queries] x [128 passages] x 768, fp32 accumulators in TMEM; the
//                       epilogue reads them back with tcgen05.ld, and appends (score, id) of the
//                       few scores >= bound to the per-query candidate list
//   (4) select (final): per query, exact radix-select + sort of the candidates with the canonical
//                       order (fp16 score desc, id asc)               -> [nq, k] scores / ids
//
// Exactness: a passage whose fp16-rounded score is below a score that k other passages reach cannot be
// in the top-k, so filtering with `>= bound` never loses a winner; ties at the bound are kept and
// resolved by id in (4).  If a candidate list overflows (massive ties, adversarial order) `status`
// is raised and the caller falls back to `atlas_b200_mips_topk_exhaustive` (chunked, no thresholds).
//
// Scan kernel roles (384 threads, persistent, 1 CTA / SM):
//   warp 0      TMA producer   : bank tile + query tile K-blocks -> 4-stage smem ring (128B swizzle)
//   warp 1      MMA issuer     : tcgen05.mma kind::f16, M=128 (x2 query halves), N=128, K=16
//   warp 2      TMEM allocator : 512 columns = 2 (double buffer) x 2 (query halves) x 128
//   warps 4-11  epilogue       : thread <-> query row (TMEM lane); threshold filter + append
#include "common.cuh"
#include "host_common.h"

#include <math.h>
#include <stdlib.h>

namespace mips {

constexpr int DIM = ATLAS_B200_EMBEDDINGS_DIM;
constexpr int BLOCK_K = 64;  // 128 bytes of 16-bit elements = one swizzle row
constexpr int K_BLOCKS = DIM / BLOCK_K;
constexpr int UMMA_K = 16;
constexpr int TILE_N = 128;   // passages per tile (UMMA N)
constexpr int HALF_M = 128;   // queries per UMMA (M)
constexpr int QBLOCK = 256;   // queries per scan launch
constexpr int STAGES = 4;
constexpr int A_HALF_BYTES = HALF_M * BLOCK_K * 2;
constexpr int B_BYTES = TILE_N * BLOCK_K * 2;
constexpr int STAGE_BYTES = 2 * A_HALF_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
constexpr int NUM_THREADS = 384;
constexpr int TMEM_COLS = 512;

static_assert(DIM % BLOCK_K == 0, "");

// ---------------------------------------------------------------------------------------------
// 16-bit float <-> order-preserving integer key (larger key = larger score; -0 canonicalised)
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t bits_to_key(uint32_t h) {
    h &= 0xFFFFu;
    if (h == 0x8000u) h = 0;
    return (h & 0x8000u) ? (~h & 0xFFFFu) : (h | 0x8000u);
}
__host__ __device__ __forceinline__ uint32_t key_to_bits(uint32_t key) {
    return (key & 0x8000u) ? (key & 0x7FFFu) : (~key & 0xFFFFu);
}

template <bool kBF16>
__device__ __forceinline__ uint32_t round_to_bits(float v) {
    if constexpr (kBF16) {
        return static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16_rn(v)));
    } else {
        return static_cast<uint32_t>(__half_as_ushort(__float2half_rn(v)));
    }
}
template <bool kBF16>
__device__ __forceinline__ float bits_to_float(uint32_t h) {
    if constexpr (kBF16) {
        return __bfloat162float(__ushort_as_bfloat16(static_cast<unsigned short>(h)));
    } else {
        return __half2float(__ushort_as_half(static_cast<unsigned short>(h)));
    }
}

__device__ __forceinline__ uint64_t pack_candidate(uint32_t key16, uint32_t id) {
    return (static_cast<uint64_t>(key16) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - id);
}

// Smallest fp32 x whose round-to-nearest-even 16-bit value is >= the value with sortable key `key16`.
template <bool kBF16>
__device__ float lower_bound_for_key(uint32_t key16) {
    const uint32_t hb = key_to_bits(key16);
    const float t = bits_to_float<kBF16>(hb);
    if (t != t) return -INFINITY;            // NaN threshold: keep everything
    if (t == -INFINITY) return -INFINITY;
    const float big = kBF16 ? 3.4028236692e38f : 65536.0f;  // first magnitude that rounds to inf
    // previous representable value (in value order, skipping the -0/+0 duplicate)
    uint32_t kp = key16 - 1;
    if (key_to_bits(kp) == 0x8000u) kp -= 1;
    float p = bits_to_float<kBF16>(key_to_bits(kp));
    if (p == -INFINITY || p != p) p = -big;
    float tt = (t == INFINITY) ? big : t;
    if (kBF16 && (fabsf(p) > 1e38f || fabsf(tt) > 1e38f)) {
        // avoid fp32 overflow of p + tt at the extreme end of the bf16 range
        return fminf(p, tt);
    }
    const float mid = 0.5f * (p + tt);  // exact: neighbours differ in one ulp of the 16-bit format
    const float r = bits_to_float<kBF16>(round_to_bits<kBF16>(mid));
    return (r >= t) ? mid : nextafterf(mid, INFINITY);
}

// ---------------------------------------------------------------------------------------------
// scan kernel
// ---------------------------------------------------------------------------------------------
struct ScanParams {
    int n_rows;       // rows in this bank shard
    int tile_begin;   // tiles visited: tile_begin + i * tile_step, i in [0, num_tiles)
    int tile_step;
    int num_tiles;
    int nq;           // valid queries in this block (<= QBLOCK)
    int n_halves;     // 1 if nq <= 128 else 2
    const float* bound;   // [QBLOCK]
    uint32_t* count;      // [QBLOCK]
    uint64_t* cand;       // [QBLOCK][capq]
    uint32_t capq;
    unsigned long long* dbg;  // optional [gridDim.x][8] cycle counters (nullptr = off)
};

template <bool kBF16>
__global__ void __launch_bounds__(NUM_THREADS, 1)
mips_scan_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_bank,
                 const ScanParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    // 1024-byte aligned tile ring (128B swizzle atoms are 1024 bytes)
    const uint32_t smem_base = (ab::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - ab::smem_u32(smem_raw));

    if (warp == 0 && lane == 0) {
        ab::tma_prefetch_desc(&tmap_q);
        ab::tma_prefetch_desc(&tmap_bank);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ab::mbar_init(&full_bar[s], 1);
            ab::mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            ab::mbar_init(&tmem_full_bar[b], 1);
            ab::mbar_init(&tmem_empty_bar[b], static_cast<uint32_t>(p.n_halves) * HALF_M);
        }
        ab::fence_barrier_init();
    }
    if (warp == 2) {
        ab::tmem_alloc<1>(&tmem_base_smem, TMEM_COLS);
    }
    ab::tc_fence_before();
    __syncthreads();
    ab::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const uint32_t tx_bytes = static_cast<uint32_t>(p.n_halves) * A_HALF_BYTES + B_BYTES;
            for (int i = blockIdx.x; i < p.num_tiles; i += gridDim.x) {
                const int tile = p.tile_begin + i * p.tile_step;
                for (int kb = 0; kb < K_BLOCKS; ++kb) {
                    ab::mbar_wait(&empty_bar[stage], phase ^ 1u, 1);
                    ab::mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
                    uint8_t* st = smem_gen + stage * STAGE_BYTES;
                    ab::tma_load_2d(&tmap_q, &full_bar[stage], st, kb * BLOCK_K, 0, ab::kEvictLast);
                    if (p.n_halves == 2)
                        ab::tma_load_2d(&tmap_q, &full_bar[stage], st + A_HALF_BYTES, kb * BLOCK_K, HALF_M,
                                        ab::kEvictLast);
                    ab::tma_load_2d(&tmap_bank, &full_bar[stage], st + 2 * A_HALF_BYTES, kb * BLOCK_K,
                                    tile * TILE_N, ab::kEvictFirst);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = ab::umma_idesc_f16(HALF_M, TILE_N, kBF16);
            uint32_t stage = 0, phase = 0;
            int it = 0;
            for (int i = blockIdx.x; i < p.num_tiles; i += gridDim.x, ++it) {
                const uint32_t buf = it & 1;
                ab::mbar_wait(&tmem_empty_bar[buf], ((it >> 1) & 1) ^ 1u, 2);
                ab::tc_fence_after();
                for (int kb = 0; kb < K_BLOCKS; ++kb) {
                    ab::mbar_wait(&full_bar[stage], phase, 3);
                    ab::tc_fence_after();
                    const uint32_t st = smem_base + stage * STAGE_BYTES;
                    for (int h = 0; h < p.n_halves; ++h) {
                        const uint32_t d_tmem = tmem_base + (buf * 2 + h) * TILE_N;
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                            const uint64_t adesc = ab::umma_desc_k_sw128(st + h * A_HALF_BYTES + k * UMMA_K * 2);
                            const uint64_t bdesc = ab::umma_desc_k_sw128(st + 2 * A_HALF_BYTES + k * UMMA_K * 2);
                            ab::umma_ss<1>(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                    }
                    ab::umma_commit(&empty_bar[stage]);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                ab::umma_commit(&tmem_full_bar[buf]);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: threshold filter + candidate append =====================
        const uint32_t half = (warp - 4) >> 2;
        const uint32_t lg = warp & 3u;  // TMEM lane group this warp may access
        if (static_cast<int>(half) < p.n_halves) {
            const uint32_t q = half * HALF_M + lg * 32 + lane;
            const float bnd = (static_cast<int>(q) < p.nq) ? p.bound[q] : INFINITY;
            uint64_t* my_cand = p.cand + static_cast<size_t>(q) * p.capq;
            int it = 0;
            for (int i = blockIdx.x; i < p.num_tiles; i += gridDim.x, ++it) {
                const uint32_t buf = it & 1;
                const int tile = p.tile_begin + i * p.tile_step;
                ab::mbar_wait(&tmem_full_bar[buf], (it >> 1) & 1, 4);
                ab::tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < TILE_N / 32; ++c) {
                    uint32_t r[32];
                    ab::tmem_ld32(tmem_base + ((lg * 32) << 16) + (buf * 2 + half) * TILE_N + c * 32, r);
                    ab::tmem_ld_wait();
                    float m = __uint_as_float(r[0]);
#pragma unroll
                    for (int j = 1; j < 32; ++j) m = fmaxf(m, __uint_as_float(r[j]));
                    if (__any_sync(0xffffffffu, m >= bnd)) {
                        const int id0 = tile * TILE_N + c * 32;
                        uint32_t npass = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            npass += (__uint_as_float(r[j]) >= bnd && id0 + j < p.n_rows) ? 1u : 0u;
                        if (npass) {
                            uint32_t pos = atomicAdd(&p.count[q], npass);
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const float v = __uint_as_float(r[j]);
                                if (v >= bnd && id0 + j < p.n_rows) {
                                    if (pos < p.capq)
                                        my_cand[pos] = pack_candidate(bits_to_key(round_to_bits<kBF16>(v)),
                                                                      static_cast<uint32_t>(id0 + j));
                                    ++pos;
                                }
                            }
                        }
                    }
                }
                ab::tc_fence_before();
                ab::mbar_arrive(&tmem_empty_bar[buf]);
            }
        }
    }

    ab::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ab::tc_fence_after();
        ab::tmem_dealloc<1>(tmem_base, TMEM_COLS);
    }
}


// ---------------------------------------------------------------------------------------------
// scan kernel v2 ("TS"): the query block stays RESIDENT ON THE SM as the tcgen05 A operand, so the
// only bytes that cross L2 -> SM are the bank rows.
//
//   128 queries x 768 fp16 = 192 KB per CTA: K-blocks 0..7 (512 K) live in TENSOR MEMORY (256 columns,
//   A-from-TMEM MMAs), K-blocks 8..11 (256 K) in 64 KB of shared memory (A-from-smem MMAs).  That split
//   leaves 256 TMEM columns for TWO 128 x 128 fp32 accumulator buffers (MMA of tile i+1 overlaps the
//   epilogue of tile i) at UMMA N = 128 — measured: below N = 128 tcgen05.mma is issue-bound at ~46-50
//   cycles per instruction, at N >= 128 it runs at the 64/128-cycle peak (tools/umma_bench.cu).
//   kPair = true : a 2-CTA cluster (cta_group::2, UMMA M=256) covers 256 queries; each CTA owns 128
//                  query rows and streams HALF of every 128-passage tile (64 full rows = 96 KB contiguous
//                  in HBM, two 3-D TMA boxes of 48 KB) through its own shared memory, so each bank byte is
//                  fetched once per SM pair.
//   kPair = false: one CTA (UMMA M=128) for <= 128 queries, four 48 KB boxes per 128-row tile.
//   smem: 64 KB resident query K-tail + 3 ring stages x 48 KB (K-major slabs [rows][128 B], 128B swizzle).
// Roles (384 threads): warp 0 TMA producer, warp 1 MMA issuer (leader CTA), warp 2 TMEM allocator,
// warps 4-11 load the query K-head into TMEM once (tcgen05.st), then run the threshold-filter epilogue
// (warp w: TMEM lane group w%4, accumulator columns 64*((w-4)/4) .. +63).
// ---------------------------------------------------------------------------------------------
constexpr int TS_TILE_N = 128;
constexpr int TS_KB_TMEM = 8;                        // K-blocks of the queries kept in TMEM
constexpr int TS_KB_SMEM = K_BLOCKS - TS_KB_TMEM;    // K-blocks of the queries kept in smem
constexpr int TS_A_COLS = TS_KB_TMEM * (BLOCK_K / 2);  // 256 TMEM columns
constexpr int TS_A_SMEM_BYTES = TS_KB_SMEM * HALF_M * 128;  // 64 KB
constexpr int TS_THREADS = 384;
constexpr int TS_EPI_WARPS = 8;

template <bool kPair>
struct TsCfg {
    static constexpr int ROWS_PER_CTA = kPair ? 64 : 128;
    static constexpr int SLAB_BYTES = ROWS_PER_CTA * 128;
    static constexpr int KB_PER_STAGE = kPair ? 6 : 3;
    static constexpr int STAGE_BYTES = KB_PER_STAGE * SLAB_BYTES;  // 48 KB
    static constexpr int STAGES_PER_TILE = K_BLOCKS / KB_PER_STAGE;
    static constexpr int STAGES = 3;
    static constexpr int UMMA_M = kPair ? 256 : 128;
    static constexpr int CTAS = kPair ? 2 : 1;
    static constexpr int SMEM_BYTES = TS_A_SMEM_BYTES + STAGES * STAGE_BYTES + 1024;
};

__device__ __forceinline__ float max32(const uint32_t (&r)[32]) {
    float t[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) t[j] = fmaxf(__uint_as_float(r[j]), __uint_as_float(r[j + 16]));
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = fmaxf(t[j], t[j + 8]);
#pragma unroll
    for (int j = 0; j < 4; ++j) t[j] = fmaxf(t[j], t[j + 4]);
    return fmaxf(fmaxf(t[0], t[1]), fmaxf(t[2], t[3]));
}

// Per-thread slot reservation in a query's candidate list: a thread claims BLOCKS of slots with one
// atomicAdd (>= 8 at a time) and fills them privately, so the L2 atomic (measured: thousands of cycles
// of latency while the bank sweep saturates HBM) is paid once or twice per thread per kernel instead of
// once per hit.  Slots of a block that stay unused are written as 0 = the smallest composite key, which
// the selection kernel can never pick while k real candidates exist.
struct SlotBlock {
    uint32_t pos;
    uint32_t left;
};

__device__ __forceinline__ void slots_release(SlotBlock& b, uint32_t capq, uint64_t* my_cand) {
    while (b.left) {
        if (b.pos < capq) my_cand[b.pos] = 0ull;
        ++b.pos;
        --b.left;
    }
}

// r[j] for a run-time j without local memory: a 5-level select tree over the 32 registers.
__device__ __forceinline__ float mux32(const uint32_t (&r)[32], int j) {
    uint32_t a[16], b[8], c[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = (j & 16) ? r[i + 16] : r[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = (j & 8) ? a[i + 8] : a[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[i + 4] : b[i];
    const uint32_t d0 = (j & 2) ? c[2] : c[0], d1 = (j & 2) ? c[3] : c[1];
    return __uint_as_float((j & 1) ? d1 : d0);
}

// Slow path of the threshold filter for one thread (= one query row) that has at least one passing
// score among 32 accumulator columns: hit bitmask by straight-line compares, then the (few) set bits are
// walked, each value fetched with the select tree above.  Only lanes with hits run this; there are no
// warp votes, no chains of dependent branches and no local-memory arrays.
template <bool kBF16>
__device__ __forceinline__ void append_hits32(const uint32_t (&r)[32], float bnd, int id0, int n_rows,
                                              uint32_t* count_q, uint32_t capq, uint64_t* my_cand, SlotBlock& blk) {
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) m |= (__uint_as_float(r[j]) >= bnd) ? (1u << j) : 0u;
    // columns past the end of the bank (zero-filled by TMA in the last tile) never count
    const int valid = n_rows - id0;
    if (valid < 32) m &= (valid <= 0) ? 0u : ((1u << valid) - 1u);
    const uint32_t npass = __popc(m);
    if (npass == 0) return;
    if (npass > blk.left) {
        slots_release(blk, capq, my_cand);
        const uint32_t want = (npass + 7u) & ~7u;
        blk.pos = atomicAdd(count_q, want);
        blk.left = want;
    }
    while (m) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        const float v = mux32(r, j);
        if (blk.pos < capq)
            my_cand[blk.pos] = pack_candidate(bits_to_key(round_to_bits<kBF16>(v)), static_cast<uint32_t>(id0 + j));
        ++blk.pos;
        --blk.left;
    }
}

template <bool kBF16, bool kPair>
__global__ void __launch_bounds__(TS_THREADS, 1)
mips_scan_ts_kernel(const __grid_constant__ CUtensorMap tmap_bank, const __grid_constant__ CUtensorMap tmap_q,
                    const uint16_t* __restrict__ qstage, const ScanParams p) {
    using C = TsCfg<kPair>;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[C::STAGES];
    __shared__ __align__(8) uint64_t empty_bar[C::STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ __align__(8) uint64_t a_tmem_bar;   // query K-head written to TMEM (all epilogue threads)
    __shared__ __align__(8) uint64_t a_smem_bar;   // query K-tail landed in smem (TMA)
    __shared__ uint32_t tmem_base_smem;

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t cta_rank = kPair ? ab::cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    const int group = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int num_groups = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    const uint32_t smem_base = (ab::smem_u32(smem_raw) + 1023u) & ~1023u;   // resident query K-tail
    const uint32_t ring_base = smem_base + TS_A_SMEM_BYTES;
    uint8_t* smem_gen = smem_raw + (smem_base - ab::smem_u32(smem_raw));
    uint8_t* ring_gen = smem_gen + TS_A_SMEM_BYTES;

    if (warp == 0 && lane == 0) {
        ab::tma_prefetch_desc(&tmap_bank);
        ab::tma_prefetch_desc(&tmap_q);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            ab::mbar_init(&full_bar[s], 1);
            ab::mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            ab::mbar_init(&tmem_full_bar[b], 1);
            ab::mbar_init(&tmem_empty_bar[b], static_cast<uint32_t>(TS_EPI_WARPS * C::CTAS));
        }
        ab::mbar_init(&a_tmem_bar, static_cast<uint32_t>(TS_EPI_WARPS * 32 * C::CTAS));
        ab::mbar_init(&a_smem_bar, 1);
        ab::fence_barrier_init();
    }
    if (warp == 2) ab::tmem_alloc<C::CTAS>(&tmem_base_smem, TMEM_COLS);
    ab::tc_fence_before();
    if constexpr (kPair) {
        ab::cluster_sync_all();
    } else {
        __syncthreads();
    }
    ab::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ===================== TMA producer (every CTA streams its own rows) =====================
        if (lane == 0) {
            // resident query K-tail: K-blocks 8..11 of this CTA's 128 query rows
            if (leader) ab::mbar_arrive_expect_tx(&a_smem_bar, TS_A_SMEM_BYTES * C::CTAS);
            for (int s = 0; s < TS_KB_SMEM; ++s) {
                if constexpr (kPair) {
                    ab::tma_load_2d_2sm(&tmap_q, &a_smem_bar, smem_gen + s * (HALF_M * 128), (TS_KB_TMEM + s) * BLOCK_K,
                                        static_cast<int>(cta_rank) * HALF_M, ab::kEvictLast);
                } else {
                    ab::tma_load_2d(&tmap_q, &a_smem_bar, smem_gen + s * (HALF_M * 128), (TS_KB_TMEM + s) * BLOCK_K, 0,
                                    ab::kEvictLast);
                }
            }
            uint32_t stage = 0, phase = 0;
            long long w_empty = 0;
            for (int i = group; i < p.num_tiles; i += num_groups) {
                const int tile = p.tile_begin + i * p.tile_step;
                const int row0 = tile * TS_TILE_N + static_cast<int>(cta_rank) * C::ROWS_PER_CTA;
                for (int st = 0; st < C::STAGES_PER_TILE; ++st) {
                    const long long t0 = clock64();
                    ab::mbar_wait(&empty_bar[stage], phase ^ 1u, 1);
                    w_empty += clock64() - t0;
                    uint8_t* dst = ring_gen + stage * C::STAGE_BYTES;
                    if constexpr (kPair) {
                        // the leader's barrier collects the bytes of BOTH CTAs' boxes
                        if (leader) ab::mbar_arrive_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
                        ab::tma_load_3d_2sm(&tmap_bank, &full_bar[stage], dst, 0, row0, st * C::KB_PER_STAGE,
                                            ab::kEvictFirst);
                    } else {
                        ab::mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
                        ab::tma_load_3d(&tmap_bank, &full_bar[stage], dst, 0, row0, st * C::KB_PER_STAGE,
                                        ab::kEvictFirst);
                    }
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
            if (p.dbg) p.dbg[blockIdx.x * 16 + 0] = w_empty;
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (lane == 0 && leader) {
            constexpr uint32_t idesc = ab::umma_idesc_f16(C::UMMA_M, TS_TILE_N, kBF16);
            const long long tstart = clock64();
            long long w_te = 0, w_full = 0;
            ab::mbar_wait(&a_tmem_bar, 0, 5);
            ab::mbar_wait(&a_smem_bar, 0, 6);
            const long long w_a = clock64() - tstart;
            ab::tc_fence_after();
            const uint64_t adesc0 = ab::umma_desc_k_sw128(smem_base);
            uint32_t stage = 0, phase = 0;
            int it = 0;
            for (int i = group; i < p.num_tiles; i += num_groups, ++it) {
                const uint32_t buf = it & 1;
                const long long t0 = clock64();
                ab::mbar_wait(&tmem_empty_bar[buf], ((it >> 1) & 1) ^ 1u, 2);
                w_te += clock64() - t0;
                ab::tc_fence_after();
                const uint32_t d_tmem = tmem_base + TS_A_COLS + buf * TS_TILE_N;
#pragma unroll
                for (int st = 0; st < C::STAGES_PER_TILE; ++st) {
                    const long long t1 = clock64();
                    ab::mbar_wait(&full_bar[stage], phase, 3);
                    w_full += clock64() - t1;
                    ab::tc_fence_after();
                    const uint64_t desc0 = ab::umma_desc_k_sw128(ring_base + stage * C::STAGE_BYTES);
#pragma unroll
                    for (int kbl = 0; kbl < C::KB_PER_STAGE; ++kbl) {
                        const int kb = st * C::KB_PER_STAGE + kbl;
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                            const uint64_t bdesc =
                                desc0 + static_cast<uint64_t>((kbl * C::SLAB_BYTES + k * UMMA_K * 2) >> 4);
                            if (kb < TS_KB_TMEM) {
                                ab::umma_ts<C::CTAS>(d_tmem, tmem_base + kb * (BLOCK_K / 2) + k * (UMMA_K / 2), bdesc,
                                                     idesc, (kb | k) != 0 ? 1u : 0u);
                            } else {
                                const uint64_t adesc =
                                    adesc0 +
                                    static_cast<uint64_t>(((kb - TS_KB_TMEM) * (HALF_M * 128) + k * UMMA_K * 2) >> 4);
                                ab::umma_ss<C::CTAS>(d_tmem, adesc, bdesc, idesc, 1u);
                            }
                        }
                    }
                    if constexpr (kPair) {
                        ab::umma_commit_2sm(&empty_bar[stage], 0x3);
                    } else {
                        ab::umma_commit(&empty_bar[stage]);
                    }
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                if constexpr (kPair) {
                    ab::umma_commit_2sm(&tmem_full_bar[buf], 0x3);
                } else {
                    ab::umma_commit(&tmem_full_bar[buf]);
                }
            }
            if (p.dbg) {
                p.dbg[blockIdx.x * 16 + 1] = w_te;
                p.dbg[blockIdx.x * 16 + 2] = w_full;
                p.dbg[blockIdx.x * 16 + 3] = w_a;
                p.dbg[blockIdx.x * 16 + 4] = clock64() - tstart;
                p.dbg[blockIdx.x * 16 + 5] = it;
            }
        }
    } else if (warp >= 4) {
        // ===================== query K-head into TMEM, then the filter epilogue =====================
        const uint32_t lg = warp & 3u;
        const uint32_t half = (warp - 4u) >> 2;                 // which 64 accumulator columns
        const uint32_t q = cta_rank * 128u + lg * 32u + lane;   // row of the staged query block
        const uint32_t lane_addr = tmem_base + ((lg * 32u) << 16);
        {
            const uint4* src = reinterpret_cast<const uint4*>(qstage + static_cast<size_t>(q) * DIM);
#pragma unroll 1
            for (int kb = static_cast<int>(half) * (TS_KB_TMEM / 2); kb < static_cast<int>(half + 1) * (TS_KB_TMEM / 2);
                 ++kb) {
                uint32_t r[32];
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    const uint4 x = __ldg(src + kb * 8 + v);
                    r[4 * v + 0] = x.x;
                    r[4 * v + 1] = x.y;
                    r[4 * v + 2] = x.z;
                    r[4 * v + 3] = x.w;
                }
                ab::tmem_st32(lane_addr + kb * (BLOCK_K / 2), r);
            }
            ab::tmem_st_wait();
            ab::tc_fence_before();
            if (leader) {
                ab::mbar_arrive(&a_tmem_bar);
            } else {
                ab::mbar_arrive_cluster(&a_tmem_bar, 0);
            }
        }
        const float bnd = (static_cast<int>(q) < p.nq) ? p.bound[q] : INFINITY;
        uint64_t* my_cand = p.cand + static_cast<size_t>(q) * p.capq;
        int it = 0;
        long long w_tf = 0, w_ld = 0, w_arr = 0, w_max = 0, w_slow = 0;
        SlotBlock blk = {0u, 0u};
        for (int i = group; i < p.num_tiles; i += num_groups, ++it) {
            const uint32_t buf = it & 1;
            const int tile = p.tile_begin + i * p.tile_step;
            const long long t0 = clock64();
            ab::mbar_wait(&tmem_full_bar[buf], (it >> 1) & 1, 4);
            const long long t1 = clock64();
            w_tf += t1 - t0;
            ab::tc_fence_after();
            const uint32_t d_addr = lane_addr + TS_A_COLS + buf * TS_TILE_N + half * 64u;
            uint32_t r0[32], r1[32];
            ab::tmem_ld32(d_addr, r0);
            ab::tmem_ld32(d_addr + 32, r1);
            ab::tmem_ld_wait();
            const long long t2 = clock64();
            w_ld += t2 - t1;
            // the accumulator columns are in registers: hand the buffer back before filtering
            ab::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) {
                    ab::mbar_arrive(&tmem_empty_bar[buf]);
                } else {
                    ab::mbar_arrive_cluster(&tmem_empty_bar[buf], 0);
                }
            }
            const long long t3 = clock64();
            w_arr += t3 - t2;
            const int id0 = tile * TS_TILE_N + static_cast<int>(half) * 64;
            // lane-divergent on purpose: only rows with a hit pay for the slow path
            const float mx0 = max32(r0), mx1 = max32(r1);
            asm volatile("" ::"f"(mx0), "f"(mx1));
            const long long t4 = clock64();
            w_max += t4 - t3;
            if (mx0 >= bnd) append_hits32<kBF16>(r0, bnd, id0, p.n_rows, &p.count[q], p.capq, my_cand, blk);
            if (mx1 >= bnd) append_hits32<kBF16>(r1, bnd, id0 + 32, p.n_rows, &p.count[q], p.capq, my_cand, blk);
            __syncwarp();
            w_slow += clock64() - t4;
        }
        slots_release(blk, p.capq, my_cand);
        if (p.dbg && warp == 4 && lane == 0) {
            p.dbg[blockIdx.x * 16 + 7] = w_ld;
            p.dbg[blockIdx.x * 16 + 8] = w_arr;
            p.dbg[blockIdx.x * 16 + 9] = w_max;
            p.dbg[blockIdx.x * 16 + 10] = w_slow;
        }
        if (p.dbg && warp == 4 && lane == 0) p.dbg[blockIdx.x * 16 + 6] = w_tf;
    }

    ab::tc_fence_before();
    if constexpr (kPair) {
        ab::cluster_sync_all();
    } else {
        __syncthreads();
    }
    if (warp == 2) {
        ab::tc_fence_after();
        ab::tmem_dealloc<C::CTAS>(tmem_base, TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------
// prep: stage one query block (zero padded to QBLOCK rows), reset counters and bounds
// ---------------------------------------------------------------------------------------------
__global__ void mips_prep_kernel(const uint16_t* __restrict__ queries, int nq, uint16_t* __restrict__ qstage,
                                 uint32_t* __restrict__ count, float* __restrict__ bound) {
    const int total = QBLOCK * DIM / 8;  // 16-byte vectors
    const uint4* src = reinterpret_cast<const uint4*>(queries);
    uint4* dst = reinterpret_cast<uint4*>(qstage);
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < total; v += gridDim.x * blockDim.x) {
        const int row = v / (DIM / 8);
        dst[v] = (row < nq) ? src[v] : make_uint4(0, 0, 0, 0);
    }
    if (blockIdx.x == 0 && threadIdx.x < QBLOCK) {
        count[threadIdx.x] = 0;
        bound[threadIdx.x] = -INFINITY;
    }
}

// ---------------------------------------------------------------------------------------------
// select: exact top-k of one query's candidate list (radix select on the 48-bit composite key,
// then a bitonic sort of the k winners).  One CTA per query.
// ---------------------------------------------------------------------------------------------
constexpr int SEL_THREADS = 256;
constexpr int SEL_WRITE_BOUND = 1;  // bound[q] = lower bound of the k-th key, count[q] = 0
constexpr int SEL_EMIT = 2;         // write out_scores / out_ids
constexpr int SEL_CARRY = 4;        // cand[q][0..k) = winners, count[q] = k

template <bool kBF16>
__global__ void __launch_bounds__(SEL_THREADS)
mips_select_kernel(uint64_t* __restrict__ cand, uint32_t* __restrict__ count, uint32_t capq, int k, int mode,
                   float* __restrict__ bound, uint16_t* __restrict__ out_scores, int64_t* __restrict__ out_ids,
                   int64_t id_base, int64_t id_stride, int32_t* __restrict__ status) {
    __shared__ uint32_t hist[256];
    __shared__ uint64_t s_prefix;
    __shared__ uint32_t s_krem;
    __shared__ uint32_t s_nwin;
    __shared__ uint64_t win[ATLAS_B200_MAX_TOPK];

    const int q = blockIdx.x;
    const int t = threadIdx.x;
    uint32_t n = count[q];
    if (n > capq) {
        if (t == 0 && status) atomicExch(status, 1);
        n = capq;
    }
    uint64_t* c = cand + static_cast<size_t>(q) * capq;
    const uint32_t kk = min(static_cast<uint32_t>(k), n);
    int P = 1;
    while (P < static_cast<int>(kk)) P <<= 1;

    if (t == 0) {
        s_prefix = 0;
        s_krem = kk;
        s_nwin = 0;
    }
    for (int i = t; i < P; i += SEL_THREADS) win[i] = 0;
    __syncthreads();

    // A bound-only pass (no winners needed) stops after the two score bytes of the 48-bit key.
    const bool bound_only = (mode & SEL_WRITE_BOUND) && !(mode & (SEL_EMIT | SEL_CARRY));
    if (kk > 0) {
        uint64_t mask = 0;
        for (int b = 5; b >= (bound_only ? 4 : 0); --b) {
            hist[t] = 0;
            __syncthreads();
            const uint64_t prefix = s_prefix;
            if (b == 5) {
                // top byte of the score key: scores cluster in a handful of bins, per-key shared atomics would
                // serialise -> lanes with the same bin elect a leader that adds their population count
                for (uint32_t i0 = 0; i0 < n; i0 += SEL_THREADS) {
                    const uint32_t i = i0 + t;
                    const bool live = i < n;
                    const uint32_t bin = live ? static_cast<uint32_t>((c[i] >> 40) & 0xFFu) : 256u;
                    const uint32_t peers = __match_any_sync(0xffffffffu, bin);
                    if (live && (t & 31) == static_cast<int>(__ffs(peers)) - 1) atomicAdd(&hist[bin], __popc(peers));
                }
            } else {
                // lower bytes are close to uniform over the 256 bins: plain shared atomics, 4 keys in flight
                for (uint32_t i = t; i < n; i += SEL_THREADS) {
                    const uint64_t key = c[i];
                    if ((key & mask) == prefix) atomicAdd(&hist[(key >> (8 * b)) & 0xFFu], 1u);
                }
            }
            __syncthreads();
            if (t == 0) {
                uint32_t krem = s_krem, acc = 0;
                int d = 255;
                for (; d > 0; --d) {
                    if (acc + hist[d] >= krem) break;
                    acc += hist[d];
                }
                s_krem = krem - acc;
                s_prefix = prefix | (static_cast<uint64_t>(d) << (8 * b));
            }
            mask |= static_cast<uint64_t>(0xFFu) << (8 * b);
            __syncthreads();
        }
        const uint64_t kth = s_prefix;  // exact k-th largest key (keys are unique: ids differ)
        for (uint32_t i = t; i < n && !bound_only; i += SEL_THREADS) {
            const uint64_t key = c[i];
            if (key >= kth) {
                const uint32_t pos = atomicAdd(&s_nwin, 1u);
                if (pos < ATLAS_B200_MAX_TOPK) win[pos] = key;
            }
        }
        __syncthreads();
        // bitonic sort, descending
        for (int size = 2; size <= P && !bound_only; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int i = t; i < P / 2; i += SEL_THREADS) {
                    const int lo = 2 * i - (i & (stride - 1));
                    const int hi = lo + stride;
                    const bool desc = ((lo & size) == 0);
                    const uint64_t a = win[lo], b2 = win[hi];
                    if ((a < b2) == desc) {
                        win[lo] = b2;
                        win[hi] = a;
                    }
                }
                __syncthreads();
            }
        }
    }

    if (mode & SEL_EMIT) {
        for (int i = t; i < k; i += SEL_THREADS) {
            uint32_t hb;
            int64_t id;
            if (i < static_cast<int>(kk)) {
                const uint64_t key = win[i];
                hb = key_to_bits(static_cast<uint32_t>(key >> 32));
                id = id_base + id_stride * static_cast<int64_t>(0xFFFFFFFFu - static_cast<uint32_t>(key));
            } else {
                hb = kBF16 ? 0xFF80u : 0xFC00u;  // -inf
                id = -1;
            }
            out_scores[static_cast<size_t>(q) * k + i] = static_cast<uint16_t>(hb);
            out_ids[static_cast<size_t>(q) * k + i] = id;
        }
    }
    if (mode & SEL_CARRY) {
        __syncthreads();
        for (int i = t; i < static_cast<int>(kk); i += SEL_THREADS) c[i] = win[i];
        if (t == 0) count[q] = kk;
    }
    if (mode & SEL_WRITE_BOUND) {
        if (t == 0) {
            if (kk == static_cast<uint32_t>(k) && kk > 0) {
                // bound-only passes resolved just the 16 score bits of the k-th key (s_prefix bits 47..32)
                const uint32_t kth16 = bound_only ? static_cast<uint32_t>(s_prefix >> 32)
                                                  : static_cast<uint32_t>(win[kk - 1] >> 32);
                bound[q] = lower_bound_for_key<kBF16>(kth16);
            } else if (!(mode & SEL_CARRY)) {
                bound[q] = -INFINITY;  // refinement passes keep the previous (still valid) bound instead
            }
            if (!(mode & SEL_CARRY)) count[q] = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// merge of W sorted per-shard lists: thread per candidate, rank by counting predecessors
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool before(uint32_t ka, int64_t ia, uint32_t kb, int64_t ib) {
    return (ka > kb) || (ka == kb && ia < ib);
}

__global__ void topk_merge_kernel(const uint16_t* __restrict__ scores_in, const int64_t* __restrict__ ids_in,
                                  int64_t stride_s, int64_t stride_i, int world, int k, int q_begin,
                                  uint16_t* __restrict__ out_scores, int64_t* __restrict__ out_ids) {
    const int qo = blockIdx.x;       // output row
    const int q = q_begin + qo;      // row inside every shard list
    for (int e = threadIdx.x; e < world * k; e += blockDim.x) {
        const int w = e / k, j = e % k;
        const size_t row = static_cast<size_t>(q) * k;
        const uint16_t hb = scores_in[w * stride_s + row + j];
        const uint32_t key = bits_to_key(hb);
        const int64_t id = ids_in[w * stride_i + row + j];
        int rank = j;  // predecessors inside its own (sorted) list
        for (int w2 = 0; w2 < world; ++w2) {
            if (w2 == w) continue;
            const uint16_t* s2 = scores_in + w2 * stride_s + row;
            const int64_t* i2 = ids_in + w2 * stride_i + row;
            int lo = 0, hi = k;  // number of elements of list w2 that come before (key, id)
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (before(bits_to_key(s2[mid]), i2[mid], key, id))
                    lo = mid + 1;
                else
                    hi = mid;
            }
            rank += lo;
        }
        if (rank < k) {
            out_scores[static_cast<size_t>(qo) * k + rank] = hb;
            out_ids[static_cast<size_t>(qo) * k + rank] = id;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// casts
// ---------------------------------------------------------------------------------------------
template <bool kBF16>
__global__ void cast_f32_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int64_t count) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < count;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        dst[i] = static_cast<uint16_t>(round_to_bits<kBF16>(src[i]));
}

template <bool kBF16>
__global__ void widen_kernel(const uint16_t* __restrict__ src, float* __restrict__ dst, int64_t count) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < count;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        dst[i] = bits_to_float<kBF16>(src[i]);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct Workspace {
    uint32_t* count;
    float* bound;
    int32_t* status;
    uint16_t* qstage;
    uint64_t* cand;
    uint32_t capq;
    // host-buffer path staging
    float* q_f32;
    uint16_t* q_16;
    uint16_t* o_s16;
    float* o_s32;
    int64_t* o_ids;
};

static uint32_t capq_for(int64_t n, int k) {
    const double s = sqrt(static_cast<double>(k) * static_cast<double>(n));
    uint64_t c = static_cast<uint64_t>(4.0 * s) + 1;
    if (c < 32768) c = 32768;
    c = (c + 1023) / 1024 * 1024;
    return static_cast<uint32_t>(c);
}

static size_t carve(Workspace* w, void* base, int64_t n, int nq, int k) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = abh::align_up(off + bytes, 256);
        return o;
    };
    const uint32_t capq = capq_for(n, k);
    const size_t o_count = take(QBLOCK * sizeof(uint32_t));
    const size_t o_bound = take(QBLOCK * sizeof(float));
    const size_t o_status = take(sizeof(int32_t));
    const size_t o_qstage = take(static_cast<size_t>(QBLOCK) * DIM * 2);
    const size_t o_cand = take(static_cast<size_t>(QBLOCK) * capq * sizeof(uint64_t));
    const size_t o_qf32 = take(static_cast<size_t>(nq) * DIM * sizeof(float));
    const size_t o_q16 = take(static_cast<size_t>(nq) * DIM * 2);
    const size_t o_os16 = take(static_cast<size_t>(nq) * k * 2);
    const size_t o_os32 = take(static_cast<size_t>(nq) * k * sizeof(float));
    const size_t o_oids = take(static_cast<size_t>(nq) * k * sizeof(int64_t));
    if (w) {
        uint8_t* b = static_cast<uint8_t*>(base);
        w->count = reinterpret_cast<uint32_t*>(b + o_count);
        w->bound = reinterpret_cast<float*>(b + o_bound);
        w->status = reinterpret_cast<int32_t*>(b + o_status);
        w->qstage = reinterpret_cast<uint16_t*>(b + o_qstage);
        w->cand = reinterpret_cast<uint64_t*>(b + o_cand);
        w->capq = capq;
        w->q_f32 = reinterpret_cast<float*>(b + o_qf32);
        w->q_16 = reinterpret_cast<uint16_t*>(b + o_q16);
        w->o_s16 = reinterpret_cast<uint16_t*>(b + o_os16);
        w->o_s32 = reinterpret_cast<float*>(b + o_os32);
        w->o_ids = reinterpret_cast<int64_t*>(b + o_oids);
    }
    return off;
}

template <bool kBF16>
static int launch_scan(const CUtensorMap& tq, const CUtensorMap& tb, const ScanParams& p, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        AB_CUDA_CHECK(cudaFuncSetAttribute(mips_scan_kernel<kBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           SMEM_BYTES));
        attr_set = true;
    }
    if (p.num_tiles <= 0) return ATLAS_B200_OK;
    const int grid = p.num_tiles < abh::num_sms() ? p.num_tiles : abh::num_sms();
    mips_scan_kernel<kBF16><<<grid, NUM_THREADS, SMEM_BYTES, s>>>(tq, tb, p);
    abh::count_launch();
    AB_CUDA_CHECK(cudaGetLastError());
    return ATLAS_B200_OK;
}

// 0 = SS kernel (queries streamed through smem), 1 = TS kernel (queries resident in TMEM)
static int g_kernel_mode = -1;
static unsigned long long* g_dbg = nullptr;  // device buffer for wait-cycle counters of the main scan
static int kernel_mode() {
    if (g_kernel_mode < 0) {
        const char* e = getenv("ATLAS_B200_MIPS_KERNEL");
        g_kernel_mode = (e && (e[0] == 's' || e[0] == 'S')) ? 0 : 1;
    }
    return g_kernel_mode;
}

template <bool kBF16, bool kPair>
static int launch_scan_ts(const CUtensorMap& tb, const CUtensorMap& tq, const uint16_t* qstage, const ScanParams& p,
                          cudaStream_t s) {
    using C = TsCfg<kPair>;
    static bool attr_set = false;
    if (!attr_set) {
        AB_CUDA_CHECK(cudaFuncSetAttribute(mips_scan_ts_kernel<kBF16, kPair>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_set = true;
    }
    if (p.num_tiles <= 0) return ATLAS_B200_OK;
    const int max_groups = abh::num_sms() / C::CTAS;
    const int groups = p.num_tiles < max_groups ? p.num_tiles : max_groups;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(groups * C::CTAS));
    cfg.blockDim = dim3(TS_THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C::CTAS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    AB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, mips_scan_ts_kernel<kBF16, kPair>, tb, tq, qstage, p));
    abh::count_launch();
    return ATLAS_B200_OK;
}

static int check_common(const void* bank, int64_t n, int64_t ld, const void* queries, int nq, int k) {
    AB_REQUIRE(nq >= 0 && k >= 1, "nq must be >= 0 and k >= 1 (nq=%d k=%d)", nq, k);
    AB_REQUIRE(k <= ATLAS_B200_MAX_TOPK, "k=%d exceeds ATLAS_B200_MAX_TOPK=%d", k, ATLAS_B200_MAX_TOPK);
    AB_REQUIRE(n >= 1 && n < (1ll << 31) - TILE_N, "bank rows n=%lld out of range", (long long)n);
    // torch.topk raises when k > n (src/index.py:118)
    AB_REQUIRE(k <= n, "selected index k out of range: k=%d > n=%lld", k, (long long)n);
    AB_REQUIRE(ld >= DIM && ld % 8 == 0, "row stride ld=%lld must be >= 768 and a multiple of 8", (long long)ld);
    AB_REQUIRE(bank != nullptr && (reinterpret_cast<uintptr_t>(bank) & 15u) == 0, "bank must be 16-byte aligned");
    AB_REQUIRE(nq == 0 || (queries != nullptr && (reinterpret_cast<uintptr_t>(queries) & 15u) == 0),
               "queries must be 16-byte aligned");
    return ATLAS_B200_OK;
}

template <bool kBF16>
static int topk_impl(const void* bank, int64_t n, int64_t ld, const void* queries, int nq, int k, void* out_scores,
                     int64_t* out_ids, int64_t id_base, int64_t id_stride, int32_t* status, void* workspace,
                     size_t workspace_bytes, cudaStream_t s, bool exhaustive) {
    int rc = check_common(bank, n, ld, queries, nq, k);
    if (rc) return rc;
    if (status) AB_CUDA_CHECK(cudaMemsetAsync(status, 0, sizeof(int32_t), s));
    if (nq == 0) return ATLAS_B200_OK;
    Workspace w;
    const size_t need = carve(&w, workspace, n, nq, k);
    if (workspace == nullptr || workspace_bytes < need) {
        abh::set_error("workspace too small: have %zu need %zu", workspace_bytes, need);
        return ATLAS_B200_EWORKSPACE;
    }
    const bool ts = kernel_mode() == 1;
    const int tile_n = ts ? TS_TILE_N : TILE_N;
    // tensor maps: SS = query block + 128-row bank boxes; TS = 64-row (single CTA) / 32-row (CTA pair) boxes
    CUtensorMap tq, tb, tb_pair;
    rc = abh::make_tmap_2d_16bit(&tq, w.qstage, QBLOCK, DIM, DIM, HALF_M, BLOCK_K, kBF16);
    if (rc) return rc;
    if (!ts) {
        rc = abh::make_tmap_2d_16bit(&tb, bank, static_cast<uint64_t>(n), DIM, static_cast<uint64_t>(ld), TILE_N,
                                     BLOCK_K, kBF16);
        if (rc) return rc;
    } else {
        rc = abh::make_tmap_kslabs_16bit(&tb, bank, static_cast<uint64_t>(n), DIM, static_cast<uint64_t>(ld),
                                         TsCfg<false>::ROWS_PER_CTA, TsCfg<false>::KB_PER_STAGE, kBF16);
        if (rc) return rc;
        rc = abh::make_tmap_kslabs_16bit(&tb_pair, bank, static_cast<uint64_t>(n), DIM, static_cast<uint64_t>(ld),
                                         TsCfg<true>::ROWS_PER_CTA, TsCfg<true>::KB_PER_STAGE, kBF16);
        if (rc) return rc;
    }

    const int total_tiles = static_cast<int>((n + tile_n - 1) / tile_n);
    const uint16_t* q16 = static_cast<const uint16_t*>(queries);
    uint16_t* os = static_cast<uint16_t*>(out_scores);

    for (int q0 = 0; q0 < nq; q0 += QBLOCK) {
        const int nqb = (nq - q0 < QBLOCK) ? (nq - q0) : QBLOCK;
        mips_prep_kernel<<<64, 256, 0, s>>>(q16 + static_cast<size_t>(q0) * DIM, nqb, w.qstage, w.count, w.bound);
        abh::count_launch();
        ScanParams p;
        p.n_rows = static_cast<int>(n);
        p.nq = nqb;
        p.n_halves = nqb > HALF_M ? 2 : 1;
        p.bound = w.bound;
        p.count = w.count;
        p.cand = w.cand;
        p.capq = w.capq;
        p.dbg = nullptr;
        uint16_t* os_b = os + static_cast<size_t>(q0) * k;
        int64_t* oi_b = out_ids + static_cast<size_t>(q0) * k;
        auto scan = [&](const ScanParams& sp) -> int {
            if (!ts) return launch_scan<kBF16>(tq, tb, sp, s);
            if (sp.n_halves == 2) return launch_scan_ts<kBF16, true>(tb_pair, tq, w.qstage, sp, s);
            return launch_scan_ts<kBF16, false>(tb, tq, w.qstage, sp, s);
        };

        if (!exhaustive) {
            int first_tile = 0;
            if (n > static_cast<int64_t>(w.capq)) {
                // (1)+(2): first bound from a strided sample of ~2*sqrt(k*n) rows (every score kept)
                // Sample size: the bound-only select costs O(m) per query and the sweep that follows keeps ~k*n_next/m
                // candidates per query.  With the refinement sweep over the first tenth of the bank (below) the next part
                // is n/10 rows, so m = sqrt(k*n/10) balances the two selects (4 Mi rows, k = 40: m = 4096 instead of the
                // 26 k of the one-stage bound 2*sqrt(k*n): the 256-query threshold select drops from ~220 us to ~35 us);
                // without a refinement stage the one-stage optimum is kept.
                double m_two_stage = sqrt(static_cast<double>(k) * static_cast<double>(n) / 10.0);
                if (m_two_stage < 4096) m_two_stage = 4096;
                const bool refine = (total_tiles / 10) * static_cast<double>(tile_n) >= 8.0 * m_two_stage;
                double m = refine ? m_two_stage : 2.0 * sqrt(static_cast<double>(k) * static_cast<double>(n));
                if (m < 4096) m = 4096;
                if (m > w.capq / 2) m = w.capq / 2;
                int sample_tiles = static_cast<int>(m / tile_n);
                if (sample_tiles < 1) sample_tiles = 1;
                if (sample_tiles > total_tiles) sample_tiles = total_tiles;
                p.tile_begin = 0;
                p.tile_step = total_tiles / sample_tiles;
                p.num_tiles = sample_tiles;
                rc = scan(p);
                if (rc) return rc;
                mips_select_kernel<kBF16><<<nqb, SEL_THREADS, 0, s>>>(w.cand, w.count, w.capq, k, SEL_WRITE_BOUND,
                                                                     w.bound, nullptr, nullptr, 0, 0, nullptr);
                abh::count_launch();
                // (2b) refinement: sweep the first ~1/10 of the bank with that bound, then tighten the bound to
                // the k-th best seen so far (its winners stay in the lists) before sweeping the other 9/10.
                // Expected survivors per query: k*n_A/m in this part, k*n_B/n_A (~9k) in the rest.
                const int tiles_a = total_tiles / 10;
                if (tiles_a * static_cast<int64_t>(tile_n) >= 8 * static_cast<int64_t>(m)) {
                    p.tile_begin = 0;
                    p.tile_step = 1;
                    p.num_tiles = tiles_a;
                    abh::prof_begin(s, abh::PROF_SCAN);
                    rc = scan(p);
                    abh::prof_end(s, abh::PROF_SCAN, static_cast<double>(tiles_a) * tile_n * DIM * 2.0);
                    if (rc) return rc;
                    mips_select_kernel<kBF16><<<nqb, SEL_THREADS, 0, s>>>(w.cand, w.count, w.capq, k,
                                                                         SEL_WRITE_BOUND | SEL_CARRY, w.bound, nullptr,
                                                                         nullptr, 0, 0, status);
                    abh::count_launch();
                    first_tile = tiles_a;
                }
            }
            // (3)+(4): the main sweep and the final selection
            p.tile_begin = first_tile;
            p.tile_step = 1;
            p.num_tiles = total_tiles - first_tile;
            p.dbg = g_dbg;
            abh::prof_begin(s, abh::PROF_SCAN);
            rc = scan(p);
            abh::prof_end(s, abh::PROF_SCAN, static_cast<double>(total_tiles - first_tile) * tile_n * DIM * 2.0);
            p.dbg = nullptr;
            if (rc) return rc;
            mips_select_kernel<kBF16><<<nqb, SEL_THREADS, 0, s>>>(w.cand, w.count, w.capq, k, SEL_EMIT, w.bound, os_b,
                                                                 oi_b, id_base, id_stride, status);
            abh::count_launch();
        } else {
            // chunked exact scan: every score of a chunk is a candidate, winners are carried forward
            int chunk_tiles = static_cast<int>((w.capq - static_cast<uint32_t>(k)) / tile_n);
            if (chunk_tiles < 1) chunk_tiles = 1;
            for (int t0 = 0; t0 < total_tiles; t0 += chunk_tiles) {
                const bool last = t0 + chunk_tiles >= total_tiles;
                p.tile_begin = t0;
                p.tile_step = 1;
                p.num_tiles = last ? total_tiles - t0 : chunk_tiles;
                rc = scan(p);
                if (rc) return rc;
                mips_select_kernel<kBF16><<<nqb, SEL_THREADS, 0, s>>>(w.cand, w.count, w.capq, k,
                                                                     last ? SEL_EMIT : SEL_CARRY, w.bound, os_b, oi_b,
                                                                     id_base, id_stride, nullptr);
                abh::count_launch();
            }
        }
    }
    AB_CUDA_CHECK(cudaGetLastError());
    return ATLAS_B200_OK;
}

}  // namespace mips

extern "C" {

void atlas_b200_mips_set_kernel(int32_t mode) { mips::g_kernel_mode = mode ? 1 : 0; }
void atlas_b200_mips_set_debug_counters(void* device_u64_buffer) {
    mips::g_dbg = static_cast<unsigned long long*>(device_u64_buffer);
}

size_t atlas_b200_mips_workspace_bytes(int64_t n, int32_t nq, int32_t k) {
    if (n < 1) n = 1;
    if (nq < 0) nq = 0;
    if (k < 1) k = 1;
    return mips::carve(nullptr, nullptr, n, nq, k);
}

int atlas_b200_mips_topk(const void* bank, int64_t n, int64_t ld, int32_t is_bf16, const void* queries, int32_t nq,
                         int32_t k, void* out_scores, int64_t* out_ids, int64_t id_base, int64_t id_stride,
                         int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (is_bf16)
        return mips::topk_impl<true>(bank, n, ld, queries, nq, k, out_scores, out_ids, id_base, id_stride, status,
                                     workspace, workspace_bytes, s, false);
    return mips::topk_impl<false>(bank, n, ld, queries, nq, k, out_scores, out_ids, id_base, id_stride, status,
                                  workspace, workspace_bytes, s, false);
}

int atlas_b200_mips_topk_exhaustive(const void* bank, int64_t n, int64_t ld, int32_t is_bf16, const void* queries,
                                    int32_t nq, int32_t k, void* out_scores, int64_t* out_ids, int64_t id_base,
                                    int64_t id_stride, void* workspace, size_t workspace_bytes, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (is_bf16)
        return mips::topk_impl<true>(bank, n, ld, queries, nq, k, out_scores, out_ids, id_base, id_stride, nullptr,
                                     workspace, workspace_bytes, s, true);
    return mips::topk_impl<false>(bank, n, ld, queries, nq, k, out_scores, out_ids, id_base, id_stride, nullptr,
                                  workspace, workspace_bytes, s, true);
}

int atlas_b200_topk_merge(const void* scores_in, const int64_t* ids_in, int64_t shard_stride_scores,
                          int64_t shard_stride_ids, int32_t is_bf16, int32_t world, int32_t nq_total, int32_t k,
                          int32_t q_begin, int32_t nq_out, void* out_scores, int64_t* out_ids, void* stream) {
    (void)is_bf16;  // the sortable-key transform is identical for both 16-bit formats
    AB_REQUIRE(world >= 1 && k >= 1 && nq_total >= 0 && nq_out >= 0 && q_begin >= 0 && q_begin + nq_out <= nq_total,
               "bad merge arguments (world=%d nq_total=%d k=%d q_begin=%d nq_out=%d)", world, nq_total, k, q_begin,
               nq_out);
    if (nq_out == 0) return ATLAS_B200_OK;
    mips::topk_merge_kernel<<<nq_out, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint16_t*>(scores_in), ids_in, shard_stride_scores, shard_stride_ids, world, k, q_begin,
        static_cast<uint16_t*>(out_scores), out_ids);
    abh::count_launch();
    AB_CUDA_CHECK(cudaGetLastError());
    return ATLAS_B200_OK;
}

int atlas_b200_cast_f32(const float* src, void* dst, int64_t count, int32_t to_bf16, void* stream) {
    if (count <= 0) return ATLAS_B200_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int64_t blocks = (count + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    if (to_bf16)
        mips::cast_f32_kernel<true><<<static_cast<int>(blocks), 256, 0, s>>>(src, static_cast<uint16_t*>(dst), count);
    else
        mips::cast_f32_kernel<false><<<static_cast<int>(blocks), 256, 0, s>>>(src, static_cast<uint16_t*>(dst), count);
    abh::count_launch();
    AB_CUDA_CHECK(cudaGetLastError());
    return ATLAS_B200_OK;
}

int atlas_b200_search_host(const void* bank, int64_t n, int64_t ld, int32_t is_bf16, const float* queries_host,
                           int32_t nq, int32_t k, float* out_scores_host, int64_t* out_ids_host, int64_t id_base,
                           int64_t id_stride, void* workspace, size_t workspace_bytes, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc = mips::check_common(bank, n, ld, reinterpret_cast<const void*>(16), nq, k);
    if (rc) return rc;
    if (nq == 0) return ATLAS_B200_OK;
    mips::Workspace w;
    const size_t need = mips::carve(&w, workspace, n, nq, k);
    if (workspace == nullptr || workspace_bytes < need) {
        abh::set_error("workspace too small: have %zu need %zu", workspace_bytes, need);
        return ATLAS_B200_EWORKSPACE;
    }
    const int64_t qcount = static_cast<int64_t>(nq) * mips::DIM;
    AB_CUDA_CHECK(cudaMemcpyAsync(w.q_f32, queries_host, qcount * sizeof(float), cudaMemcpyHostToDevice, s));
    rc = atlas_b200_cast_f32(w.q_f32, w.q_16, qcount, is_bf16, stream);
    if (rc) return rc;
    rc = atlas_b200_mips_topk(bank, n, ld, is_bf16, w.q_16, nq, k, w.o_s16, w.o_ids, id_base, id_stride, w.status,
                              workspace, workspace_bytes, stream);
    if (rc) return rc;
    int32_t st = 0;
    AB_CUDA_CHECK(cudaMemcpyAsync(&st, w.status, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    const int64_t ocount = static_cast<int64_t>(nq) * k;
    int blocks = static_cast<int>((ocount + 255) / 256);
    if (blocks > 1024) blocks = 1024;
    if (is_bf16)
        mips::widen_kernel<true><<<blocks, 256, 0, s>>>(w.o_s16, w.o_s32, ocount);
    else
        mips::widen_kernel<false><<<blocks, 256, 0, s>>>(w.o_s16, w.o_s32, ocount);
    abh::count_launch();
    AB_CUDA_CHECK(cudaMemcpyAsync(out_scores_host, w.o_s32, ocount * sizeof(float), cudaMemcpyDeviceToHost, s));
    AB_CUDA_CHECK(cudaMemcpyAsync(out_ids_host, w.o_ids, ocount * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    AB_CUDA_CHECK(cudaStreamSynchronize(s));
    if (st != 0) {
        rc = atlas_b200_mips_topk_exhaustive(bank, n, ld, is_bf16, w.q_16, nq, k, w.o_s16, w.o_ids, id_base, id_stride,
                                             workspace, workspace_bytes, stream);
        if (rc) return rc;
        if (is_bf16)
            mips::widen_kernel<true><<<blocks, 256, 0, s>>>(w.o_s16, w.o_s32, ocount);
        else
            mips::widen_kernel<false><<<blocks, 256, 0, s>>>(w.o_s16, w.o_s32, ocount);
        abh::count_launch();
        AB_CUDA_CHECK(cudaMemcpyAsync(out_scores_host, w.o_s32, ocount * sizeof(float), cudaMemcpyDeviceToHost, s));
        AB_CUDA_CHECK(cudaMemcpyAsync(out_ids_host, w.o_ids, ocount * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        AB_CUDA_CHECK(cudaStreamSynchronize(s));
    }
    return ATLAS_B200_OK;
}

}  // extern "C"
