// attention_lanes_kernel: the encoder-shaped attention of Contriever and FiD (>= 2 query tiles per (segment, head), <= 576
// keys) - the second-generation forward kernel.  Same math and the same call (atlas_b200_attention_ex) as csrc/attention.cu:
//     O[b, i, h, :] = softmax_j( scale * Q.K + rel_bias[h, j - i] + key_mask[b, j] (+ causal) ) V
// replacing BertSelfAttention.forward (src/modeling_bert.py:328-366) and T5Attention.forward (src/modeling_t5.py:478-524).
//
// What limited the first generation (profiles/r01_attention_v2_v3.md: tensor pipe 13 %, 14 k cycles per 128 x 384 tile against
// a MUFU floor of 3.1 k): ONE query tile in flight per CTA, so every softmax phase waited for the P.V -> S round trip of its
// own tile; 8 softmax warps could not hide the tcgen05.ld / shared-memory latencies; and ~7 issue slots per score element.
// Here:
//   * THREE independent "lanes" per CTA, one per 128-row query tile of the current (segment, head) - all 384 rows of a
//     FiD passage are in flight at once.  Every lane has its own MMA-issuer thread, its own softmax warpgroup (one thread
//     per query row: no cross-thread reductions, no named barriers) and its own barriers; the lanes only share the K / V
//     stream.  While lane A waits for its P.V(j) -> S(j+1) hand-over, lanes B and C keep the MUFU / FMA pipes busy.
//   * K and V stream ONCE per (segment, head) through a ring of 64-key chunks (8 KB, TMA, 128B swizzle) shared by the
//     lanes (a stage is released when every query tile of the item has consumed it): L2 -> SM traffic stays at
//     (K + V) per item, the next item's chunks prefetch into the freed stages.
//   * Online softmax over 64-key blocks with a LAZY reference maximum: block 0 fixes m; a later block only triggers a
//     rescale of the 64-column O accumulator when its maximum exceeds m by more than 8 (log2 units), otherwise the stale m
//     is kept (probabilities up to 2^8, exact in fp32 / harmless in 16 bits) - the common case costs nothing.
//   * TMEM (480 of 512 columns): per lane S (64 fp32 columns) | P (32 columns of packed 16-bit pairs) | O (64).  P has its OWN
//     columns, so S(j+1) = Q K_{j+1}^T is issued as soon as the softmax threads have READ S(j) into registers (`s_read`
//     barrier) and runs on the tensor pipe while they do the arithmetic of block j: a softmax warp never waits for an MMA
//     round trip.  The first version (csrc/attention_lanes96.cu: P written over S, S(j+1) only after P.V(j) had retired)
//     spent 24 % of its softmax-warp samples in that wait (profiles/r02_attention_lanes.md).  The issuer never waits for
//     its own MMAs either: tcgen05.mma executes in issue order, so P.V(j) -> S(j+2) -> ... need no completion barrier
//     between them; only the softmax threads wait for P.V(j-1) before they overwrite P (it has long retired by then).
//   * The score pipeline per element: 1/4 LDS.128 (relative-position bias from FOUR alignment-shifted copies of the
//     [2L - 1] table, so that every thread reads its diagonal run with 16-byte loads) + 1/2 FFMA2 + 1/2 FMNMX3 + 1/2 FADD2
//     + 1 MUFU.EX2 + 1/2 FADD2 + 1/2 F2FP: 3.75 issue slots (packed f32x2 arithmetic and the 3-input max are sm_100
//     instructions), the t values stay in registers between the two passes (setmaxnreg moves registers from the four
//     helper warps to the twelve softmax warps).
// Packed layout (Params::seg_tile, the padding-compacted FiD encoder): an item has 1 - 3 query tiles; tile qt of the item runs on
// lane (rot + qt) % 3, rot = the number of tiles of the CTA's earlier items - the lanes stay evenly loaded and drift apart by up
// to the ring's depth.  A lane without a tile in an item passes its K / V chunks (waits for the chunk, releases it).
// Warp roles (512 threads): warp 0 = tables (all lanes) + TMA producer (lane 0); warps 1-3 = MMA issuers of lanes 0-2;
// warps 4-7 / 8-11 / 12-15 = softmax + output of lanes 0 / 1 / 2 (thread = query row = TMEM lane).
#include "common.cuh"
#include "host_common.h"

#include <math.h>
#include <stdlib.h>

#include <type_traits>

namespace attn4 {

constexpr int D = 64;
constexpr int BQ = 128;                 // query rows per tile / lane
constexpr int BK = 64;                  // keys per block (UMMA N of S, K extent of P.V)
constexpr int LANES = 3;
constexpr int RING = 18;                // K / V chunk stages (8 KB each): a whole 576-key item, or 384 keys + prefetch
constexpr int STAGE_BYTES = BK * D * 2;
constexpr int Q_BYTES = BQ * D * 2;
constexpr int MAX_BLOCKS = 9;           // <= 576 keys
constexpr int MAXK = MAX_BLOCKS * BK;
constexpr int THREADS = 512;
constexpr int SM_THREADS = 128 * LANES;
constexpr int LANE_COLS = 160;          // TMEM columns per lane: S at +0 (64), P at +64 (32), O at +96 (64)
constexpr int P_OFF = 64;
constexpr int O_OFF = 96;
constexpr int TMEM_COLS = 512;
constexpr int CPLEN = 1152;             // floats per shifted bias copy (>= MAXK + 512 + 4), multiple of 32
constexpr int CPSTRIDE = CPLEN + 8;     // +32 bytes per copy: the four copies sit in different 16-byte bank groups
constexpr int SMEM_BYTES = 1024 + LANES * Q_BYTES + RING * STAGE_BYTES;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_GAP = 8.0f;     // lazy rescale threshold, log2 units

struct Params {
    int B, H, Lq, Lk;
    int q_col0, k_col0, v_col0;
    uint16_t* O;
    int64_t ldo;
    const float* add_mask;      // [B, Lk] additive key mask or nullptr
    const float* bias_delta;    // [H, Lq + Lk - 1] or nullptr
    float scale;
    float causal_value;
    float* lse_out;             // [B, H, Lq] or nullptr
    const uint8_t* blk_live;    // [B, ceil(Lk / 64)] or nullptr: 0 = every key of the 64-key block is masked out (additive mask
                                // <= -5000: softmax weight exactly 0 in fp32) -> the block is neither loaded nor computed.
                                // FiD passages are padded to text_maxlength; a segment keeps at least one live block.
    const int32_t* seg_tile;    // nullptr, or [B * ceil(Lk / 64)] (atlas_b200_segment_tile_scan): PACKED layout - segment b's rows
                                // are the 64-row tiles it keeps (blk_live: a prefix of its tiles), stored back to back from row
                                // 64 * seg_tile[b * nb] of q / k / v / out; query tiles past the kept rows do not exist.  Needs
                                // Lq == Lk <= 384 and blk_live.
    const int32_t* seg_work;    // optional with seg_tile: [B + 1] exclusive prefix of the segments' work (kept key blocks x query
                                // tiles): the CTAs split the head-major item list by work instead of by item count
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// packed fp32 pairs (FFMA2 / FADD2): d = a * s + b,  d = a + b
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float s, float b0, float b1) {
    uint64_t a, b, c, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(s), "f"(s));
    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(b0), "f"(b1));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    uint64_t a, b, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("add.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

// MN-major operand tile (rows of 128 bytes = 64 head dims, 128B swizzle): 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1024 >> 4) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// Pass 1 on one 32-key chunk, in place: r[jj] <- t = S * scale2 (+ bias2[j]) (+ mask2[j]) in the log2 domain; returns
// max(mx, chunk maximum).  `pb` points at this thread's diagonal run of its alignment copy of the bias table (16-byte
// aligned for j0 % 4 == 0); the key mask is the same for every row (broadcast 16-byte loads).
template <bool kBias, bool kMask>
__device__ __forceinline__ float chunk_scores(uint32_t (&r)[32], float scale2, const float* __restrict__ pb,
                                              const float* __restrict__ mask2, float mx) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if constexpr (kBias) {
            const float4 b = *reinterpret_cast<const float4*>(pb + 4 * q);
            a0 = b.x, a1 = b.y, a2 = b.z, a3 = b.w;
        }
        if constexpr (kMask) {
            const float4 m = *reinterpret_cast<const float4*>(mask2 + 4 * q);
            if constexpr (kBias) {
                fadd2(a0, a1, a0, a1, m.x, m.y);
                fadd2(a2, a3, a2, a3, m.z, m.w);
            } else {
                a0 = m.x, a1 = m.y, a2 = m.z, a3 = m.w;
            }
        }
        float t0, t1, t2, t3;
        ffma2(t0, t1, __uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), scale2, a0, a1);
        ffma2(t2, t3, __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]), scale2, a2, a3);
        r[4 * q] = __float_as_uint(t0);
        r[4 * q + 1] = __float_as_uint(t1);
        r[4 * q + 2] = __float_as_uint(t2);
        r[4 * q + 3] = __float_as_uint(t3);
        mx = fmax3(mx, t0, t1);
        mx = fmax3(mx, t2, t3);
    }
    return mx;
}

// Pass 2 on one chunk: p = 2^(t - m), packed to 16 bits; the partial sums accumulate as a pair.
template <bool kBF16>
__device__ __forceinline__ void chunk_probs(const uint32_t (&r)[32], uint32_t (&pk)[16], float neg_m, float& s0, float& s1) {
#pragma unroll
    for (int jj = 0; jj < 32; jj += 2) {
        float d0, d1;
        fadd2(d0, d1, __uint_as_float(r[jj]), __uint_as_float(r[jj + 1]), neg_m, neg_m);
        const float e0 = ex2_approx(d0), e1 = ex2_approx(d1);
        fadd2(s0, s1, s0, s1, e0, e1);
        pk[jj >> 1] = ab::pack2_rn<kBF16>(e0, e1);
    }
}

template <bool kBF16>
__global__ void __launch_bounds__(THREADS, 1)
attention_lanes_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                       const __grid_constant__ CUtensorMap tmap_v, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t q_full[LANES], q_empty[LANES], s_full[LANES], s_read[LANES], p_ready[LANES], pv_done[LANES];
    __shared__ __align__(8) uint64_t kv_full[RING], kv_empty[RING], tab_full[2], tab_empty[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ int s_mask_flag[2];                              // this item's key mask has a non-zero entry
    __shared__ uint32_t s_live[2];                              // this item's live key blocks (bit j = block j)
    __shared__ __align__(16) float s_bias[4 * CPSTRIDE];        // 4 alignment-shifted copies of (bias (+ causal)) * log2e
    __shared__ __align__(16) float s_mask[2][MAXK];             // additive key mask * log2e, -inf beyond Lk

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t smem_base = (ab::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - ab::smem_u32(smem_raw));
    uint8_t* sQ = smem_gen;
    uint8_t* sRing = smem_gen + LANES * Q_BYTES;
    const uint32_t aQ = smem_base, aRing = smem_base + LANES * Q_BYTES;

    const int nb = (p.Lk + BK - 1) / BK;                        // key blocks per tile (<= MAX_BLOCKS, host-checked)
    const int lk_pad = nb * BK;
    const int n_qt = (p.Lq + BQ - 1) / BQ;                      // query tiles per item; tile t belongs to lane t % 3
    const int n_items = p.B * p.H;
    __shared__ int s_range[2];
    const bool has_bias = (p.bias_delta != nullptr) || (p.causal_value != 0.f);
    const bool packed = p.seg_tile != nullptr;

    if (warp == 0 && lane == 0) {
        ab::tma_prefetch_desc(&tmap_q);
        ab::tma_prefetch_desc(&tmap_k);
        ab::tma_prefetch_desc(&tmap_v);
    }
    if (warp == 1 && lane == 0) {
        for (int l = 0; l < LANES; ++l) {
            ab::mbar_init(&q_full[l], 1);
            ab::mbar_init(&q_empty[l], 1);
            ab::mbar_init(&s_full[l], 1);
            ab::mbar_init(&s_read[l], 128);
            ab::mbar_init(&p_ready[l], 128);
            ab::mbar_init(&pv_done[l], 1);
        }
        for (int s = 0; s < RING; ++s) {
            ab::mbar_init(&kv_full[s], 1);
            // every query tile of the item consumes the chunk once; packed layout: every lane releases it once (use or pass)
            ab::mbar_init(&kv_empty[s], static_cast<uint32_t>(packed ? LANES : n_qt));
        }
        for (int i = 0; i < 2; ++i) {
            ab::mbar_init(&tab_full[i], 32);
            ab::mbar_init(&tab_empty[i], SM_THREADS);
        }
        ab::fence_barrier_init();
    }
    if (warp == 2) ab::tmem_alloc<1>(&tmem_base_smem, TMEM_COLS);
    if (warp == 3 && lane < 2) {
        // contiguous item range of this CTA, head-major (item = h * B + b): the bias tables change at most twice per CTA.
        // Packed layout: the items differ in size (1 - 6 kept key blocks x 1 - 3 query tiles) - split by work, not by count.
        const int64_t c = static_cast<int64_t>(blockIdx.x) + lane;
        int bound = static_cast<int>(c * n_items / gridDim.x);
        if (p.seg_work != nullptr) {
            const int64_t w_seg = __ldg(p.seg_work + p.B);                       // work of one head (> 0)
            const int64_t t = c * (w_seg * p.H) / gridDim.x;
            const int h = static_cast<int>(t / w_seg);
            const int r = static_cast<int>(t % w_seg);
            int lo = 0, hi = p.B;                                                // first b with prefix[b + 1] > r
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (__ldg(p.seg_work + mid + 1) > r) hi = mid;
                else lo = mid + 1;
            }
            bound = c >= gridDim.x ? n_items : min(h * p.B + lo, n_items);
        }
        s_range[lane] = bound;
    }
    ab::tc_fence_before();
    __syncthreads();
    ab::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    const int it_begin = s_range[0], it_end = s_range[1];

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 0) {
            // ===================== tables (all lanes) + TMA producer (lane 0), one item ahead of the consumers ==========
            uint32_t chunk_ctr = 0;
            uint32_t q_ctr[LANES] = {0, 0, 0};
            int prev_h = -1;
            int item_it = 0;
            int rot = 0;
            for (int it = it_begin; it < it_end; ++it, ++item_it) {
                const int h = it / p.B, b = it % p.B;
                const int buf = item_it & 1;
                // the item's mask row in two batches of independent loads (two L2 latencies instead of MAXK / 32 dependent ones;
                // this warp runs on 40 registers)
                ab::mbar_wait_nocall(&tab_empty[buf], ((item_it >> 1) & 1) ^ 1u);
                bool nonzero = false;
                constexpr int HALF = MAXK / 64;
#pragma unroll 1
                for (int hf = 0; hf < 2; ++hf) {
                    float mv[HALF];
#pragma unroll
                    for (int q = 0; q < HALF; ++q) {
                        const int j = static_cast<int>(lane) + 32 * (hf * HALF + q);
                        mv[q] = (p.add_mask != nullptr && j < p.Lk) ? __ldg(p.add_mask + static_cast<size_t>(b) * p.Lk + j) : 0.f;
                    }
#pragma unroll
                    for (int q = 0; q < HALF; ++q) {
                        const int j = static_cast<int>(lane) + 32 * (hf * HALF + q);
                        if (j < lk_pad) {
                            const float v = j < p.Lk ? mv[q] * LOG2E : -INFINITY;
                            nonzero |= (j < p.Lk) && (v != 0.f);
                            s_mask[buf][j] = v;
                        }
                    }
                }
                nonzero = __any_sync(0xffffffffu, nonzero);
                if (lane == 0) s_mask_flag[buf] = nonzero ? 1 : 0;
                uint32_t live = (1u << nb) - 1u;
                if (p.blk_live != nullptr) {
                    const bool lv = static_cast<int>(lane) < nb && __ldg(p.blk_live + static_cast<size_t>(b) * nb + lane) != 0;
                    const uint32_t m = __ballot_sync(0xffffffffu, lv);
                    if (m != 0u) live = m;
                }
                if (lane == 0) s_live[buf] = live;
                if (has_bias && h != prev_h) {
                    // the single bias buffer is shared by consecutive items of one head: before rewriting it every reader of
                    // the previous item must be done (rare: the head changes at most twice per CTA)
                    if (item_it > 0) ab::mbar_wait_nocall(&tab_empty[buf ^ 1], ((item_it - 1) >> 1) & 1u);
                    const int n_valid = p.Lq + p.Lk - 1;
                    for (int r = 0; r < 4; ++r)
                        for (int x = static_cast<int>(lane); x < CPLEN; x += 32) {
                            const int d = x + r;
                            float v = 0.f;
                            if (d < n_valid) {
                                v = p.bias_delta ? p.bias_delta[static_cast<size_t>(h) * n_valid + d] : 0.f;
                                if (p.causal_value != 0.f && d > p.Lq - 1) v += p.causal_value;   // j > i
                            }
                            s_bias[r * CPSTRIDE + x] = v * LOG2E;
                        }
                    prev_h = h;
                }
                ab::mbar_arrive(&tab_full[buf]);
                // rows of the item: packed layout = the segment's kept tiles, back to back
                const int item_qt = packed ? (__popc(live) + 1) / 2 : n_qt;
                const int q_row0 = packed ? 64 * __ldg(p.seg_tile + static_cast<size_t>(b) * nb) : b * p.Lq;
                const int k_row0 = packed ? q_row0 : b * p.Lk;
                if (lane == 0) {
                    auto load_q = [&](int qt) {
                        const int l = (rot + qt) % LANES;
                        ab::mbar_wait_nocall(&q_empty[l], (q_ctr[l] & 1u) ^ 1u);
                        ab::mbar_arrive_expect_tx(&q_full[l], Q_BYTES);
                        ab::tma_load_2d(&tmap_q, &q_full[l], sQ + l * Q_BYTES, p.q_col0 + h * D, q_row0 + qt * BQ,
                                        ab::kEvictFirst);
                        ++q_ctr[l];
                    };
                    auto load_chunk = [&](const CUtensorMap* map, int col0, int j) {
                        const uint32_t st = chunk_ctr % RING;
                        ab::mbar_wait_nocall(&kv_empty[st], ((chunk_ctr / RING) & 1u) ^ 1u);
                        ab::mbar_arrive_expect_tx(&kv_full[st], STAGE_BYTES);
                        ab::tma_load_2d(map, &kv_full[st], sRing + st * STAGE_BYTES, col0 + h * D, k_row0 + j * BK,
                                        ab::kEvictNormal);
                        ++chunk_ctr;
                    };
                    for (int qt = 0; qt < item_qt && qt < LANES; ++qt) load_q(qt);
                    for (uint32_t lm = live; lm != 0u; lm &= lm - 1u) {      // live key blocks only, in order
                        const int j = __ffs(lm) - 1;
                        load_chunk(&tmap_k, p.k_col0, j);
                        load_chunk(&tmap_v, p.v_col0, j);
                    }
                    for (int qt = LANES; qt < item_qt; ++qt) load_q(qt);
                }
                if (packed) rot = (rot + item_qt) % LANES;
                __syncwarp();
            }
        } else if (lane == 0) {
            // ===================== MMA issuer of lane `l`: S(j) = Q K_j^T, then O (+)= P(j) V_j =====================
            const int l = static_cast<int>(warp) - 1;
            constexpr uint32_t idesc_s = ab::umma_idesc_f16(BQ, BK, kBF16);
            constexpr uint32_t idesc_o = ab::umma_idesc_f16(BQ, D, kBF16) | (1u << 16);   // B = V rows, MN-major
            const uint32_t s_tmem = tmem_base + l * LANE_COLS;
            const uint32_t o_tmem = s_tmem + O_OFF;
            const uint64_t qdesc = ab::umma_desc_k_sw128(aQ + l * Q_BYTES);
            const uint32_t p_tmem = s_tmem + P_OFF;
            uint32_t tile_ctr = 0, blk_ctr = 0;     // blk_ctr: blocks whose P.V has been issued
            uint32_t s_ctr = 0;                     // blocks whose S has been issued (runs one ahead inside a tile)
            int item_it = 0;
            // S(n): needs the K chunk and the softmax threads' read of S(n - 1) (they hold it in registers: s_read)
            auto issue_s = [&](uint32_t ck, bool last_of_tile) {
                const uint32_t sk = ck % RING;
                ab::mbar_wait_nocall(&kv_full[sk], (ck / RING) & 1u);
                if (s_ctr > 0) ab::mbar_wait_nocall(&s_read[l], (s_ctr - 1) & 1u);
                ab::tc_fence_after();
                const uint64_t kdesc = ab::umma_desc_k_sw128(aRing + sk * STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < D / 16; ++k)
                    ab::umma_ss<1>(s_tmem, qdesc + ((k * 32) >> 4), kdesc + ((k * 32) >> 4), idesc_s, k != 0 ? 1u : 0u);
                ab::umma_commit(&kv_empty[sk]);
                if (last_of_tile) ab::umma_commit(&q_empty[l]);
                ab::umma_commit(&s_full[l]);
                ++s_ctr;
            };
            uint32_t chunk_base = 0;                 // K / V chunks of the items before this one (2 per live block)
            int rot = 0;
            for (int it = it_begin; it < it_end; ++it, ++item_it) {
                // the item's live key blocks: published by warp 0 with the tables (read once, at the item's start)
                ab::mbar_wait_nocall(&tab_full[item_it & 1], (item_it >> 1) & 1u);
                const uint32_t live = s_live[item_it & 1];
                const int nl = __popc(live);
                const int item_qt = packed ? (nl + 1) / 2 : n_qt;
                const int qt0 = packed ? (l + LANES - rot) % LANES : l;
                if (packed) {
                    rot = (rot + item_qt) % LANES;
                    if (qt0 >= item_qt) {
                        // no tile of this item on this lane: pass its chunks (the stage is released when all lanes have)
                        for (uint32_t c = 0; c < 2u * static_cast<uint32_t>(nl); ++c) {
                            const uint32_t ck = chunk_base + c;
                            ab::mbar_wait_nocall(&kv_full[ck % RING], (ck / RING) & 1u);
                            ab::mbar_arrive(&kv_empty[ck % RING]);
                        }
                    }
                }
                for (int qt = qt0; qt < item_qt; qt += LANES, ++tile_ctr) {
                    ab::mbar_wait_nocall(&q_full[l], tile_ctr & 1u);
                    issue_s(chunk_base, nl == 1);
                    for (int j = 0; j < nl; ++j, ++blk_ctr) {           // j = ordinal among the live blocks
                        // S(j + 1) goes to the tensor pipe while the softmax threads work on block j
                        if (j + 1 < nl) issue_s(chunk_base + 2u * (j + 1), j + 2 == nl);
                        const uint32_t cv = chunk_base + 2u * j + 1u;
                        const uint32_t sv = cv % RING;
                        ab::mbar_wait_nocall(&kv_full[sv], (cv / RING) & 1u);
                        ab::mbar_wait_nocall(&p_ready[l], blk_ctr & 1u);
                        ab::tc_fence_after();
                        const uint64_t vdesc = umma_desc_mn_sw128(aRing + sv * STAGE_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            ab::umma_ts<1>(o_tmem, p_tmem + k * 8, vdesc + static_cast<uint64_t>((k * 2048) >> 4), idesc_o,
                                           (j != 0 || k != 0) ? 1u : 0u);
                        ab::umma_commit(&kv_empty[sv]);
                        ab::umma_commit(&pv_done[l]);
                    }
                }
                chunk_base += 2u * static_cast<uint32_t>(nl);
            }
        }
    } else {
        // ===================== softmax + output: lane l = warps 4+4l .. 7+4l, thread = query row = TMEM lane ===========
        asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
        const int l = static_cast<int>(warp - 4u) >> 2;
        const uint32_t quad = warp & 3u;
        const int row = static_cast<int>(quad * 32u + lane);
        const uint32_t s_addr = tmem_base + ((quad * 32u) << 16) + static_cast<uint32_t>(l * LANE_COLS);
        const uint32_t p_addr = s_addr + P_OFF;
        const uint32_t o_addr = s_addr + O_OFF;
        const float scale2 = p.scale * LOG2E;
        const bool partial_last = (p.Lk != lk_pad);
        uint32_t blk_ctr = 0;

        // The output of a finished tile is written one block LATE, inside block 0 of the lane's next tile (after that block's
        // pass 1): by then P.V(last) of the finished tile has long retired, so its completion latency and the TMEM read of O
        // hide behind useful work instead of stalling the warpgroup at every tile end (17 % of the softmax-warp samples in
        // the ncu capture of the previous version, profiles/r02_attention_lanes64_ncu_source_top.csv).  O stays intact until
        // this lane's next P.V(0), which is issued only after the p_ready(0) that follows the write-out.
        bool pend = false;
        int pend_b = 0, pend_h = 0, pend_i = 0, pend_lq = 0;
        int64_t pend_row0 = 0;                           // first output row of the pending tile's item
        float pend_m = 0.f, pend_sum = 1.f;
        auto emit_pending = [&]() {
            const float inv = 1.0f / pend_sum;
            if (p.lse_out != nullptr && pend_i < pend_lq)
                p.lse_out[(static_cast<size_t>(pend_b) * p.H + pend_h) * p.Lq + pend_i] = pend_m * (1.0f / LOG2E) + __logf(pend_sum);
#pragma unroll 1
            for (int cc = 0; cc < D / 32; ++cc) {
                uint32_t ro[32];
                ab::tmem_ld32(o_addr + cc * 32, ro);
                ab::tmem_ld_wait();
                if (pend_i < pend_lq) {
                    uint4* dst = reinterpret_cast<uint4*>(p.O + (pend_row0 + pend_i) * p.ldo + pend_h * D + cc * 32);
#pragma unroll
                    for (int v4 = 0; v4 < 4; ++v4)
                        dst[v4] = make_uint4(
                            ab::pack2_rn<kBF16>(__uint_as_float(ro[8 * v4]) * inv, __uint_as_float(ro[8 * v4 + 1]) * inv),
                            ab::pack2_rn<kBF16>(__uint_as_float(ro[8 * v4 + 2]) * inv, __uint_as_float(ro[8 * v4 + 3]) * inv),
                            ab::pack2_rn<kBF16>(__uint_as_float(ro[8 * v4 + 4]) * inv, __uint_as_float(ro[8 * v4 + 5]) * inv),
                            ab::pack2_rn<kBF16>(__uint_as_float(ro[8 * v4 + 6]) * inv, __uint_as_float(ro[8 * v4 + 7]) * inv));
                }
            }
            pend = false;
        };

        // all key blocks of one query tile; kBias / kMaskAll are tile-uniform, the last block of a ragged segment always
        // applies the mask (its pad keys carry -inf)
        uint32_t tile_live = 0;                        // live key blocks of the current item
        auto run_tile = [&](auto bias_tag, auto mask_tag, const float* pb_row, const float* mask2, float& m_ref, float& sum0,
                            float& sum1) {
            constexpr bool kBias = decltype(bias_tag)::value;
            constexpr bool kMaskAll = decltype(mask_tag)::value;
            bool first = true;
            for (uint32_t lm = tile_live; lm != 0u; lm &= lm - 1u, ++blk_ctr, first = false) {
                const int j = __ffs(lm) - 1;                           // key block index (live blocks only)
                const bool use_mask = kMaskAll || (j == nb - 1 && partial_last);
                ab::mbar_wait_nocall(&s_full[l], blk_ctr & 1u);
                ab::tc_fence_after();
                uint32_t r[BK / 32][32];
#pragma unroll
                for (int c = 0; c < BK / 32; ++c) ab::tmem_ld32(s_addr + c * 32, r[c]);
                ab::tmem_ld_wait();
                ab::tc_fence_before();
                ab::mbar_arrive(&s_read[l]);             // S(j) is in registers: the issuer may overwrite it with S(j + 1)
                // ---- pass 1: t = scaled score + bias (+ mask), block maximum ----
                float mb = -INFINITY;
                const float* pb = pb_row + j * BK;
                const float* mk = mask2 + j * BK;
                if (use_mask) {
#pragma unroll
                    for (int c = 0; c < BK / 32; ++c) mb = chunk_scores<kBias, true>(r[c], scale2, pb + c * 32, mk + c * 32, mb);
                } else {
#pragma unroll
                    for (int c = 0; c < BK / 32; ++c) mb = chunk_scores<kBias, false>(r[c], scale2, pb + c * 32, mk + c * 32, mb);
                }
                // P(j - 1) must have been consumed before P(j) overwrites its columns, and O must be quiescent for a rescale:
                // P.V(j - 1) was issued when this block started and has normally retired by now
                if (blk_ctr > 0) {
                    ab::mbar_wait_nocall(&pv_done[l], (blk_ctr - 1) & 1u);
                    ab::tc_fence_after();
                }
                // ---- lazy reference maximum ----
                if (first) {
                    if (pend) emit_pending();            // the previous tile's output (its last P.V retired above)
                    m_ref = (mb == -INFINITY) ? 0.f : mb;
                } else {
                    const bool need = mb > m_ref + RESCALE_GAP;
                    if (__any_sync(0xffffffffu, need)) {
                        const float f = need ? ex2_approx(m_ref - mb) : 1.0f;
                        if (need) m_ref = mb;
                        sum0 *= f;
                        sum1 *= f;
#pragma unroll 1
                        for (int cc = 0; cc < D / 16; ++cc) {
                            uint32_t ro[16];
                            tmem_ld16(o_addr + cc * 16, ro);
                            ab::tmem_ld_wait();
#pragma unroll
                            for (int e = 0; e < 16; ++e) ro[e] = __float_as_uint(__uint_as_float(ro[e]) * f);
                            tmem_st16(o_addr + cc * 16, ro);
                        }
                    }
                }
                // ---- pass 2: p = 2^(t - m_ref), packed into the lane's P columns ----
                const float neg_m = -m_ref;
#pragma unroll
                for (int c = 0; c < BK / 32; ++c) {
                    uint32_t pk[16];
                    chunk_probs<kBF16>(r[c], pk, neg_m, sum0, sum1);
                    tmem_st16(p_addr + c * 16, pk);
                }
                ab::tmem_st_wait();
                ab::tc_fence_before();
                ab::mbar_arrive(&p_ready[l]);
            }
        };

        int item_it = 0;
        int rot = 0;
        for (int it = it_begin; it < it_end; ++it, ++item_it) {
            const int h = it / p.B, b = it % p.B;
            const int buf = item_it & 1;
            ab::mbar_wait_nocall(&tab_full[buf], (item_it >> 1) & 1u);
            const float* mask2 = s_mask[buf];
            const bool item_mask = s_mask_flag[buf] != 0;
            tile_live = s_live[buf];
            const int nl = __popc(tile_live);
            const int item_qt = packed ? (nl + 1) / 2 : n_qt;
            const int qt0 = packed ? (l + LANES - rot) % LANES : l;
            if (packed) rot = (rot + item_qt) % LANES;
            const int item_lq = packed ? nl * BK : p.Lq;
            const int64_t item_row0 = packed ? 64ll * __ldg(p.seg_tile + static_cast<size_t>(b) * nb) : static_cast<int64_t>(b) * p.Lq;
            for (int qt = qt0; qt < item_qt; qt += LANES) {
                const int i = qt * BQ + row;                          // query position inside the segment
                const int off = max(p.Lq - 1 - i, 0);                 // bias index = j + off (clamped for pad rows)
                const float* pb_row = s_bias + (off & 3) * CPSTRIDE + (off & ~3);
                float m_ref = 0.f, sum0 = 0.f, sum1 = 0.f;
                if (has_bias) {
                    if (item_mask) run_tile(std::true_type{}, std::true_type{}, pb_row, mask2, m_ref, sum0, sum1);
                    else run_tile(std::true_type{}, std::false_type{}, pb_row, mask2, m_ref, sum0, sum1);
                } else {
                    if (item_mask) run_tile(std::false_type{}, std::true_type{}, pb_row, mask2, m_ref, sum0, sum1);
                    else run_tile(std::false_type{}, std::false_type{}, pb_row, mask2, m_ref, sum0, sum1);
                }
                // ---- output: deferred into block 0 of this lane's next tile (emit_pending) ----
                pend = true;
                pend_b = b, pend_h = h, pend_i = i, pend_m = m_ref, pend_sum = sum0 + sum1;
                pend_lq = item_lq, pend_row0 = item_row0;
            }
            ab::mbar_arrive(&tab_empty[buf]);
        }
        if (pend) {                                      // the lane's last tile
            ab::mbar_wait_nocall(&pv_done[l], (blk_ctr - 1) & 1u);
            ab::tc_fence_after();
            emit_pending();
        }
    }

    ab::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ab::tc_fence_after();
        ab::tmem_dealloc<1>(tmem_base, TMEM_COLS);
    }
}

}  // namespace attn4

// Launched by atlas_b200_attention_ex (csrc/attention.cu) for >= 2 query tiles per item; returns ATLAS_B200_OK or an error.
int atlas_b200_attention_lanes_launch(const void* q, int64_t ldq, int32_t q_col0, const void* k, int64_t ldk, int32_t k_col0,
                                      const void* v, int64_t ldv, int32_t v_col0, void* out, int64_t ldo,
                                      const float* add_mask, const float* bias_delta, int32_t B, int32_t H, int32_t Lq,
                                      int32_t Lk, float scale, float causal_value, float* lse_out, const uint8_t* blk_live,
                                      const int32_t* seg_tile, const int32_t* seg_work, int32_t is_bf16, cudaStream_t s) {
    using namespace attn4;
    AB_REQUIRE(Lk <= MAXK && Lq <= 512, "attention_lanes: Lq <= 512 and Lk <= %d", MAXK);
    AB_REQUIRE(seg_tile == nullptr || (blk_live != nullptr && Lq == Lk && Lk % BK == 0 && Lq <= LANES * BQ && lse_out == nullptr),
               "attention_lanes: the packed layout needs blk_live, Lq == Lk <= %d, Lk %% 64 == 0, no lse", LANES * BQ);
    CUtensorMap tq, tk, tv;
    int rc = abh::make_tmap_2d_16bit(&tq, q, static_cast<uint64_t>(B) * Lq, static_cast<uint64_t>(q_col0 + H * D),
                                     static_cast<uint64_t>(ldq), BQ, D, is_bf16 != 0);
    if (rc) return rc;
    rc = abh::make_tmap_2d_16bit(&tk, k, static_cast<uint64_t>(B) * Lk, static_cast<uint64_t>(k_col0 + H * D),
                                 static_cast<uint64_t>(ldk), BK, D, is_bf16 != 0);
    if (rc) return rc;
    rc = abh::make_tmap_2d_16bit(&tv, v, static_cast<uint64_t>(B) * Lk, static_cast<uint64_t>(v_col0 + H * D),
                                 static_cast<uint64_t>(ldv), BK, D, is_bf16 != 0);
    if (rc) return rc;
    Params p;
    p.B = B, p.H = H, p.Lq = Lq, p.Lk = Lk;
    p.q_col0 = q_col0, p.k_col0 = k_col0, p.v_col0 = v_col0;
    p.O = static_cast<uint16_t*>(out);
    p.ldo = ldo;
    p.add_mask = add_mask;
    p.bias_delta = bias_delta;
    p.scale = scale;
    p.causal_value = causal_value;
    p.lse_out = lse_out;
    p.blk_live = blk_live;
    p.seg_tile = seg_tile;
    p.seg_work = seg_tile != nullptr ? seg_work : nullptr;
    const int items = B * H;
    const int grid = items < abh::num_sms() ? items : abh::num_sms();
    static bool attr_set[2] = {false, false};
    if (is_bf16) {
        if (!attr_set[0]) {
            AB_CUDA_CHECK(cudaFuncSetAttribute(attention_lanes_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            attr_set[0] = true;
        }
        attention_lanes_kernel<true><<<grid, THREADS, SMEM_BYTES, s>>>(tq, tk, tv, p);
    } else {
        if (!attr_set[1]) {
            AB_CUDA_CHECK(cudaFuncSetAttribute(attention_lanes_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            attr_set[1] = true;
        }
        attention_lanes_kernel<false><<<grid, THREADS, SMEM_BYTES, s>>>(tq, tk, tv, p);
    }
    AB_CUDA_CHECK(cudaGetLastError());
    return ATLAS_B200_OK;
}
