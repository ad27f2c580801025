// EXPERIMENT (ATLAS_B200_ATTN_LANES=2): 48-key blocks with a double-buffered S per lane.  Round-2 visit C measured it no faster
// than the 96-key kernel of attention_lanes.cu and wrong at the +48 column buffer; kept for the A/B record, not dispatched by default.
// attention_lanes2_kernel: the encoder-shaped attention of Contriever and FiD (>= 2 query tiles per (segment, head), <= 576
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
//   * K and V stream ONCE per (segment, head) through a ring of 48-key chunks (6 KB, TMA, 128B swizzle) shared by the
//     lanes (a stage is released when every query tile of the item has consumed it): L2 -> SM traffic stays at
//     (K + V) per item, the next item's chunks prefetch into the freed stages.
//   * S is DOUBLE-BUFFERED per lane: the issuer runs S(j+1) = Q K_{j+1}^T into the other buffer while the softmax
//     warpgroup works on block j, so a softmax warp never waits for an MMA round trip (v1 of this kernel had one 96-key
//     S buffer per lane and spent half of its time in those waits: profiles/r02_attention_lanes.md).
//   * Online softmax over 48-key blocks with a LAZY reference maximum: block 0 fixes m; a later block only triggers a
//     rescale of the 64-column O accumulator when its maximum exceeds m by more than 8 (log2 units), otherwise the stale m
//     is kept (probabilities up to 2^8, exact in fp32 / harmless in 16 bits) - the common case costs nothing.
//   * TMEM (480 of 512 columns): per lane S0 | S1 (48 fp32 columns each, overwritten in place by the packed 16-bit P) | O (64).
//   * The score pipeline per element: 1/4 LDS.128 (relative-position bias from FOUR alignment-shifted copies of the
//     [2L - 1] table, so that every thread reads its diagonal run with 16-byte loads) + 1/2 FFMA2 + 1/2 FMNMX3 + 1/2 FADD2
//     + 1 MUFU.EX2 + 1/2 FADD2 + 1/2 F2FP: 3.75 issue slots (packed f32x2 arithmetic and the 3-input max are sm_100
//     instructions), the t values stay in registers between the two passes (setmaxnreg moves registers from the four
//     helper warps to the twelve softmax warps).
// Warp roles (512 threads): warp 0 = tables (all lanes) + TMA producer (lane 0); warps 1-3 = MMA issuers of lanes 0-2;
// warps 4-7 / 8-11 / 12-15 = softmax + output of lanes 0 / 1 / 2 (thread = query row = TMEM lane).
#include "common.cuh"
#include "host_common.h"

#include <math.h>
#include <stdlib.h>

#include <type_traits>

namespace attn5 {

constexpr int D = 64;
constexpr int BQ = 128;                 // query rows per tile / lane
constexpr int BK = 48;                  // keys per block (UMMA N of S, K extent of P.V)
constexpr int LANES = 3;
constexpr int RING = 24;                // K / V chunk stages (6 KB each): a whole 576-key item, or 384 keys + prefetch
constexpr int STAGE_BYTES = BK * D * 2;
constexpr int Q_BYTES = BQ * D * 2;
constexpr int MAX_BLOCKS = 12;          // <= 576 keys
constexpr int MAXK = MAX_BLOCKS * BK;
constexpr int THREADS = 512;
constexpr int SM_THREADS = 128 * LANES;
constexpr int LANE_COLS = 160;          // TMEM columns per lane: S / P buffer 0 at +0 (48), buffer 1 at +48, O at +96 (64)
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
    int debug;                  // ATLAS_B200_ATTN_DEBUG bit mask (timing experiments only; results are wrong when set):
                                //   1 = no bias / mask loads, 2 = no MUFU (p = t), 4 = no output stores, 8 = no S load
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

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

// Pass 1 on the 48 keys of a block, in place: t = S * scale2 (+ bias2[j]) (+ mask2[j]) in the log2 domain; returns the block
// maximum.  `pb` points at this thread's diagonal run of its alignment copy of the bias table (16-byte aligned: block and
// unit offsets are multiples of 4); the key mask is the same for every row (broadcast 16-byte loads).  All twelve bias
// loads are issued before the first use and four independent maxima are kept: the block is latency-, not issue-bound.
template <bool kBias, bool kMask>
__device__ __forceinline__ float block_scores(uint32_t (&r)[BK], float scale2, const float* __restrict__ pb,
                                              const float* __restrict__ mask2) {
    float4 add[BK / 4];
#pragma unroll
    for (int q = 0; q < BK / 4; ++q) {
        if constexpr (kBias) add[q] = *reinterpret_cast<const float4*>(pb + 4 * q);
        else add[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if constexpr (kMask) {
#pragma unroll
        for (int q = 0; q < BK / 4; ++q) {
            const float4 m = *reinterpret_cast<const float4*>(mask2 + 4 * q);
            if constexpr (kBias) {
                fadd2(add[q].x, add[q].y, add[q].x, add[q].y, m.x, m.y);
                fadd2(add[q].z, add[q].w, add[q].z, add[q].w, m.z, m.w);
            } else {
                add[q] = m;
            }
        }
    }
    float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int q = 0; q < BK / 4; ++q) {
        float t0, t1, t2, t3;
        ffma2(t0, t1, __uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), scale2, add[q].x, add[q].y);
        ffma2(t2, t3, __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]), scale2, add[q].z, add[q].w);
        r[4 * q] = __float_as_uint(t0);
        r[4 * q + 1] = __float_as_uint(t1);
        r[4 * q + 2] = __float_as_uint(t2);
        r[4 * q + 3] = __float_as_uint(t3);
        mx[q & 1] = fmax3(mx[q & 1], t0, t1);
        mx[2 + (q & 1)] = fmax3(mx[2 + (q & 1)], t2, t3);
    }
    return fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
}

// Pass 2 on the block: p = 2^(t - m) packed to 16 bits (24 words); the row sum accumulates in two independent pairs.
template <bool kBF16>
__device__ __forceinline__ void block_probs(uint32_t (&r)[BK], uint32_t (&pk)[BK / 2], float neg_m, float (&sum)[4]) {
#pragma unroll
    for (int jj = 0; jj < BK; jj += 2) {
        float d0, d1;
        fadd2(d0, d1, __uint_as_float(r[jj]), __uint_as_float(r[jj + 1]), neg_m, neg_m);
        r[jj] = __float_as_uint(d0);
        r[jj + 1] = __float_as_uint(d1);
    }
#pragma unroll
    for (int jj = 0; jj < BK; ++jj) r[jj] = __float_as_uint(ex2_approx(__uint_as_float(r[jj])));
#pragma unroll
    for (int jj = 0; jj < BK; jj += 2) {
        const int a = (jj >> 1) & 1;
        fadd2(sum[2 * a], sum[2 * a + 1], sum[2 * a], sum[2 * a + 1], __uint_as_float(r[jj]), __uint_as_float(r[jj + 1]));
        pk[jj >> 1] = ab::pack2_rn<kBF16>(__uint_as_float(r[jj]), __uint_as_float(r[jj + 1]));
    }
}

template <bool kBF16>
__global__ void __launch_bounds__(THREADS, 1)
attention_lanes2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                       const __grid_constant__ CUtensorMap tmap_v, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t q_full[LANES], q_empty[LANES], s_full[LANES][2], p_ready[LANES][2], pv_done[LANES];
    __shared__ __align__(8) uint64_t kv_full[RING], kv_empty[RING], tab_full[2], tab_empty[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ int s_mask_flag[2];                              // this item's key mask has a non-zero entry
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
    // contiguous item range of this CTA, head-major (item = h * B + b): the bias tables change at most twice per CTA
    const int it_begin = static_cast<int>(static_cast<int64_t>(blockIdx.x) * n_items / gridDim.x);
    const int it_end = static_cast<int>(static_cast<int64_t>(blockIdx.x + 1) * n_items / gridDim.x);
    const bool has_bias = (p.bias_delta != nullptr) || (p.causal_value != 0.f);

    if (warp == 0 && lane == 0) {
        ab::tma_prefetch_desc(&tmap_q);
        ab::tma_prefetch_desc(&tmap_k);
        ab::tma_prefetch_desc(&tmap_v);
    }
    if (warp == 1 && lane == 0) {
        for (int l = 0; l < LANES; ++l) {
            ab::mbar_init(&q_full[l], 1);
            ab::mbar_init(&q_empty[l], 1);
            ab::mbar_init(&s_full[l][0], 1);
            ab::mbar_init(&s_full[l][1], 1);
            ab::mbar_init(&p_ready[l][0], 128);
            ab::mbar_init(&p_ready[l][1], 128);
            ab::mbar_init(&pv_done[l], 1);
        }
        for (int s = 0; s < RING; ++s) {
            ab::mbar_init(&kv_full[s], 1);
            ab::mbar_init(&kv_empty[s], static_cast<uint32_t>(n_qt));   // every query tile of the item consumes the chunk once
        }
        for (int i = 0; i < 2; ++i) {
            ab::mbar_init(&tab_full[i], 32);
            ab::mbar_init(&tab_empty[i], SM_THREADS);
        }
        ab::fence_barrier_init();
    }
    if (warp == 2) ab::tmem_alloc<1>(&tmem_base_smem, TMEM_COLS);
    ab::tc_fence_before();
    __syncthreads();
    ab::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 0) {
            // ===================== tables (all lanes) + TMA producer (lane 0), one item ahead of the consumers ==========
            uint32_t chunk_ctr = 0;
            uint32_t q_ctr[LANES] = {0, 0, 0};
            int prev_h = -1;
            int item_it = 0;
            for (int it = it_begin; it < it_end; ++it, ++item_it) {
                const int h = it / p.B, b = it % p.B;
                const int buf = item_it & 1;
                ab::mbar_wait_nocall(&tab_empty[buf], ((item_it >> 1) & 1) ^ 1u);
                bool nonzero = false;
                for (int j = static_cast<int>(lane); j < lk_pad; j += 32) {
                    float v = -INFINITY;
                    if (j < p.Lk) {
                        v = p.add_mask ? p.add_mask[static_cast<size_t>(b) * p.Lk + j] * LOG2E : 0.f;
                        nonzero |= (v != 0.f);
                    }
                    s_mask[buf][j] = v;
                }
                nonzero = __any_sync(0xffffffffu, nonzero);
                if (lane == 0) s_mask_flag[buf] = nonzero ? 1 : 0;
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
                if (lane == 0) {
                    auto load_q = [&](int qt) {
                        const int l = qt % LANES;
                        ab::mbar_wait_nocall(&q_empty[l], (q_ctr[l] & 1u) ^ 1u);
                        ab::mbar_arrive_expect_tx(&q_full[l], Q_BYTES);
                        ab::tma_load_2d(&tmap_q, &q_full[l], sQ + l * Q_BYTES, p.q_col0 + h * D, b * p.Lq + qt * BQ,
                                        ab::kEvictFirst);
                        ++q_ctr[l];
                    };
                    auto load_chunk = [&](const CUtensorMap* map, int col0, int j) {
                        const uint32_t st = chunk_ctr % RING;
                        ab::mbar_wait_nocall(&kv_empty[st], ((chunk_ctr / RING) & 1u) ^ 1u);
                        ab::mbar_arrive_expect_tx(&kv_full[st], STAGE_BYTES);
                        ab::tma_load_2d(map, &kv_full[st], sRing + st * STAGE_BYTES, col0 + h * D, b * p.Lk + j * BK,
                                        ab::kEvictNormal);
                        ++chunk_ctr;
                    };
                    // the first two key blocks go first (they land in stages the previous item no longer needs, so they
                    // are resident when the lanes reach this item), then the query tiles (each waits for its lane's last
                    // S MMA of the previous tile), then the rest of K / V
                    const int head_blocks = nb < 2 ? nb : 2;
                    for (int j = 0; j < head_blocks; ++j) {
                        load_chunk(&tmap_k, p.k_col0, j);
                        load_chunk(&tmap_v, p.v_col0, j);
                    }
                    for (int qt = 0; qt < n_qt && qt < LANES; ++qt) load_q(qt);
                    for (int j = head_blocks; j < nb; ++j) {
                        load_chunk(&tmap_k, p.k_col0, j);
                        load_chunk(&tmap_v, p.v_col0, j);
                    }
                    for (int qt = LANES; qt < n_qt; ++qt) load_q(qt);
                }
                __syncwarp();
            }
        } else if (lane == 0) {
            // ===================== MMA issuer of lane `l` =====================
            // Block n (global count over this lane's tiles) uses S / P buffer n % 2.  Order per tile:
            //   S(0) | S(1), P.V(0) | S(2), P.V(1) | ... | P.V(nb-1)      - S runs one block ahead of the softmax.
            // S(n) overwrites the buffer that held P(n-2): P.V(n-2) must have retired (pv_done), which it has long before
            // the softmax of block n-1 ends.
            const int l = static_cast<int>(warp) - 1;
            constexpr uint32_t idesc_s = ab::umma_idesc_f16(BQ, BK, kBF16);
            constexpr uint32_t idesc_o = ab::umma_idesc_f16(BQ, D, kBF16) | (1u << 16);   // B = V rows, MN-major
            const uint32_t lane_tmem = tmem_base + l * LANE_COLS;
            const uint32_t o_tmem = lane_tmem + O_OFF;
            const uint64_t qdesc = ab::umma_desc_k_sw128(aQ + l * Q_BYTES);
            uint32_t tile_ctr = 0, blk_ctr = 0;     // blk_ctr: blocks whose P.V has been issued
            int item_it = 0;
            // `wait_pv`: index of the P.V completion to wait for (always the LATEST one issued so far, so the parity test is
            // unambiguous); -1 = none.  S(n) needs P.V(n-2) retired; at a tile start all P.V of the previous tile are issued.
            auto issue_s = [&](uint32_t n, uint32_t ck, bool last_of_tile, int wait_pv) {
                const uint32_t sk = ck % RING;
                ab::mbar_wait_nocall(&kv_full[sk], (ck / RING) & 1u);
                if (wait_pv >= 0) ab::mbar_wait_nocall(&pv_done[l], static_cast<uint32_t>(wait_pv) & 1u);
                ab::tc_fence_after();
                const uint64_t kdesc = ab::umma_desc_k_sw128(aRing + sk * STAGE_BYTES);
                const uint32_t s_tmem = lane_tmem + (n & 1u) * BK;
#pragma unroll
                for (int k = 0; k < D / 16; ++k)
                    ab::umma_ss<1>(s_tmem, qdesc + ((k * 32) >> 4), kdesc + ((k * 32) >> 4), idesc_s, k != 0 ? 1u : 0u);
                ab::umma_commit(&kv_empty[sk]);
                if (last_of_tile) ab::umma_commit(&q_empty[l]);
                ab::umma_commit(&s_full[l][n & 1u]);
            };
            for (int it = it_begin; it < it_end; ++it, ++item_it) {
                const uint32_t chunk_base = static_cast<uint32_t>(item_it) * 2u * nb;
                for (int qt = l; qt < n_qt; qt += LANES, ++tile_ctr) {
                    ab::mbar_wait_nocall(&q_full[l], tile_ctr & 1u);
                    issue_s(blk_ctr, chunk_base, nb == 1, static_cast<int>(blk_ctr) - 1);
                    for (int j = 0; j < nb; ++j, ++blk_ctr) {
                        if (j + 1 < nb) issue_s(blk_ctr + 1, chunk_base + 2u * (j + 1), j + 2 == nb, static_cast<int>(blk_ctr) - 1);
                        const uint32_t cv = chunk_base + 2u * j + 1u;
                        const uint32_t sv = cv % RING;
                        ab::mbar_wait_nocall(&kv_full[sv], (cv / RING) & 1u);
                        ab::mbar_wait_nocall(&p_ready[l][blk_ctr & 1u], (blk_ctr >> 1) & 1u);
                        ab::tc_fence_after();
                        const uint64_t vdesc = umma_desc_mn_sw128(aRing + sv * STAGE_BYTES);
                        const uint32_t p_tmem = lane_tmem + (blk_ctr & 1u) * BK;
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            ab::umma_ts<1>(o_tmem, p_tmem + k * 8, vdesc + static_cast<uint64_t>((k * 2048) >> 4), idesc_o,
                                           (j != 0 || k != 0) ? 1u : 0u);
                        ab::umma_commit(&kv_empty[sv]);
                        ab::umma_commit(&pv_done[l]);
                    }
                }
            }
        }
    } else {
        // ===================== softmax + output: lane l = warps 4+4l .. 7+4l, thread = query row = TMEM lane ===========
        asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
        const int l = static_cast<int>(warp - 4u) >> 2;
        const uint32_t quad = warp & 3u;
        const int row = static_cast<int>(quad * 32u + lane);
        const uint32_t lane_addr = tmem_base + ((quad * 32u) << 16) + static_cast<uint32_t>(l * LANE_COLS);
        const uint32_t o_addr = lane_addr + O_OFF;
        const float scale2 = p.scale * LOG2E;
        const bool partial_last = (p.Lk != lk_pad);
        uint32_t blk_ctr = 0;

        // all key blocks of one query tile; kBias / kMaskAll are tile-uniform, the last block of a ragged segment always
        // applies the mask (its pad keys carry -inf)
        auto run_tile = [&](auto bias_tag, auto mask_tag, const float* pb_row, const float* mask2, float& m_ref, float& sum_out) {
            constexpr bool kBias = decltype(bias_tag)::value;
            constexpr bool kMaskAll = decltype(mask_tag)::value;
            float sum[4] = {0.f, 0.f, 0.f, 0.f};
            for (int j = 0; j < nb; ++j, ++blk_ctr) {
                const bool use_mask = kMaskAll || (j == nb - 1 && partial_last);
                const uint32_t s_addr = lane_addr + (blk_ctr & 1u) * BK;
                ab::mbar_wait_nocall(&s_full[l][blk_ctr & 1u], (blk_ctr >> 1) & 1u);
                ab::tc_fence_after();
                uint32_t r[BK];
                if (!(p.debug & 8)) {
                    uint32_t r0[32], r1[16];
                    ab::tmem_ld32(s_addr, r0);
                    tmem_ld16(s_addr + 32, r1);
                    ab::tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 32; ++e) r[e] = r0[e];
#pragma unroll
                    for (int e = 0; e < 16; ++e) r[32 + e] = r1[e];
                } else {
#pragma unroll
                    for (int e = 0; e < BK; ++e) r[e] = __float_as_uint(0.001f * static_cast<float>(e + row));
                }
                // ---- pass 1: t = scaled score + bias (+ mask), block maximum ----
                const float* pb = pb_row + j * BK;
                const float* mk = mask2 + j * BK;
                float mb;
                if (p.debug & 1) mb = block_scores<false, false>(r, scale2, pb, mk);
                else if (use_mask) mb = block_scores<kBias, true>(r, scale2, pb, mk);
                else mb = block_scores<kBias, false>(r, scale2, pb, mk);
                // ---- lazy reference maximum ----
                if (j == 0) {
                    m_ref = (mb == -INFINITY) ? 0.f : mb;
                } else {
                    const bool need = mb > m_ref + RESCALE_GAP;
                    if (__any_sync(0xffffffffu, need)) {
                        // O may only be touched once P.V(j-1) has retired (S runs ahead of it): rare path, explicit wait
                        ab::mbar_wait_nocall(&pv_done[l], (blk_ctr - 1) & 1u);
                        ab::tc_fence_after();
                        const float f = need ? ex2_approx(m_ref - mb) : 1.0f;
                        if (need) m_ref = mb;
#pragma unroll
                        for (int e = 0; e < 4; ++e) sum[e] *= f;
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
                // ---- pass 2: p = 2^(t - m_ref), packed P over the first half of this S buffer ----
                uint32_t pk[BK / 2];
                if (p.debug & 2) {
#pragma unroll
                    for (int e = 0; e < BK / 2; ++e) {
                        pk[e] = ab::pack2_rn<kBF16>(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
                        sum[0] += __uint_as_float(r[2 * e]);
                    }
                } else {
                    block_probs<kBF16>(r, pk, -m_ref, sum);
                }
                {
                    uint32_t p0[16], p1[8];
#pragma unroll
                    for (int e = 0; e < 16; ++e) p0[e] = pk[e];
#pragma unroll
                    for (int e = 0; e < 8; ++e) p1[e] = pk[16 + e];
                    tmem_st16(s_addr, p0);
                    tmem_st8(s_addr + 16, p1);
                }
                ab::tmem_st_wait();
                ab::tc_fence_before();
                ab::mbar_arrive(&p_ready[l][blk_ctr & 1u]);
            }
            sum_out = (sum[0] + sum[1]) + (sum[2] + sum[3]);
        };

        int item_it = 0;
        for (int it = it_begin; it < it_end; ++it, ++item_it) {
            const int h = it / p.B, b = it % p.B;
            const int buf = item_it & 1;
            ab::mbar_wait_nocall(&tab_full[buf], (item_it >> 1) & 1u);
            const float* mask2 = s_mask[buf];
            const bool item_mask = s_mask_flag[buf] != 0;
            for (int qt = l; qt < n_qt; qt += LANES) {
                const int i = qt * BQ + row;                          // query position inside the segment
                const int off = max(p.Lq - 1 - i, 0);                 // bias index = j + off (clamped for pad rows)
                const float* pb_row = s_bias + (off & 3) * CPSTRIDE + (off & ~3);
                float m_ref = 0.f, sum = 0.f;
                if (has_bias) {
                    if (item_mask) run_tile(std::true_type{}, std::true_type{}, pb_row, mask2, m_ref, sum);
                    else run_tile(std::true_type{}, std::false_type{}, pb_row, mask2, m_ref, sum);
                } else {
                    if (item_mask) run_tile(std::false_type{}, std::true_type{}, pb_row, mask2, m_ref, sum);
                    else run_tile(std::false_type{}, std::false_type{}, pb_row, mask2, m_ref, sum);
                }
                // ---- output: O / sum ----
                ab::mbar_wait_nocall(&pv_done[l], (blk_ctr - 1) & 1u);
                ab::tc_fence_after();
                const float inv = 1.0f / sum;
                if (p.lse_out != nullptr && i < p.Lq)
                    p.lse_out[(static_cast<size_t>(b) * p.H + h) * p.Lq + i] = m_ref * (1.0f / LOG2E) + __logf(sum);
#pragma unroll 1
                for (int cc = 0; cc < D / 32; ++cc) {
                    uint32_t ro[32];
                    ab::tmem_ld32(o_addr + cc * 32, ro);
                    ab::tmem_ld_wait();
                    if (i < p.Lq && !(p.debug & 4)) {
                        uint4* dst = reinterpret_cast<uint4*>(p.O + (static_cast<size_t>(b) * p.Lq + i) * p.ldo + h * D + cc * 32);
#pragma unroll
                        for (int v4 = 0; v4 < 4; ++v4)
                            dst[v4] = make_uint4(
                                ab::pack2_rn<kBF16>(__uint_as_float(ro[8 * v4]) * inv, __uint_as_float(ro[8 * v4 + 1]) * inv),
                                ab::pack2_rn<kBF16>(__uint_as_float(ro[8 * v4 + 2]) * inv, __uint_as_float(ro[8 * v4 + 3]) * inv),
                                ab::pack2_rn<kBF16>(__uint_as_float(ro[8 * v4 + 4]) * inv, __uint_as_float(ro[8 * v4 + 5]) * inv),
                                ab::pack2_rn<kBF16>(__uint_as_float(ro[8 * v4 + 6]) * inv, __uint_as_float(ro[8 * v4 + 7]) * inv));
                    }
                }
                ab::tc_fence_before();      // the next tile's first p_ready orders these O reads before P.V(0) overwrites O
            }
            ab::mbar_arrive(&tab_empty[buf]);
        }
    }

    ab::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ab::tc_fence_after();
        ab::tmem_dealloc<1>(tmem_base, TMEM_COLS);
    }
}

}  // namespace attn5

// Launched by atlas_b200_attention_ex (csrc/attention.cu) for >= 2 query tiles per item; returns ATLAS_B200_OK or an error.
int atlas_b200_attention_lanes2_launch(const void* q, int64_t ldq, int32_t q_col0, const void* k, int64_t ldk, int32_t k_col0,
                                      const void* v, int64_t ldv, int32_t v_col0, void* out, int64_t ldo,
                                      const float* add_mask, const float* bias_delta, int32_t B, int32_t H, int32_t Lq,
                                      int32_t Lk, float scale, float causal_value, float* lse_out, int32_t is_bf16,
                                      cudaStream_t s) {
    using namespace attn5;
    AB_REQUIRE(Lk <= MAXK && Lq <= 512, "attention_lanes: Lq <= 512 and Lk <= %d", MAXK);
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
    static const int dbg = getenv("ATLAS_B200_ATTN_DEBUG") ? atoi(getenv("ATLAS_B200_ATTN_DEBUG")) : 0;
    p.debug = dbg;
    const int items = B * H;
    const int grid = items < abh::num_sms() ? items : abh::num_sms();
    static bool attr_set[2] = {false, false};
    if (is_bf16) {
        if (!attr_set[0]) {
            AB_CUDA_CHECK(cudaFuncSetAttribute(attention_lanes2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            attr_set[0] = true;
        }
        attention_lanes2_kernel<true><<<grid, THREADS, SMEM_BYTES, s>>>(tq, tk, tv, p);
    } else {
        if (!attr_set[1]) {
            AB_CUDA_CHECK(cudaFuncSetAttribute(attention_lanes2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            attr_set[1] = true;
        }
        attention_lanes2_kernel<false><<<grid, THREADS, SMEM_BYTES, s>>>(tq, tk, tv, p);
    }
    AB_CUDA_CHECK(cudaGetLastError());
    return ATLAS_B200_OK;
}
