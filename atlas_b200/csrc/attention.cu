// Fused multi-head attention (head_dim 64, <= 512 keys per segment) on tcgen05 for the Contriever and
// FiD encoders and the FiD decoder's self-attention:
//     O[b, i, h, :] = softmax_j( scale * Q[b,i,h,:].K[b,j,h,:] + rel_bias[h, j-i] + key_mask[b, j] (+ causal) ) V[b, j, h, :]
// Replaces, per layer, two batched cuBLAS GEMMs + the materialised [B, H, L, L] score / probability tensors +
// the ATen softmax / mask / bias kernels of
//     BertSelfAttention.forward  src/modeling_bert.py:328-366   (scale 1/8, additive key mask, fp32 softmax)
//     T5Attention.forward        src/modeling_t5.py:478-524      (no scaling, relative-position bias + mask
//                                                                 added to the scores, fp32 softmax)
// Q, K, V are read IN PLACE from the fused projection output ([tokens, n_cols] row-major, head h at a
// column offset) through TMA boxes {64 columns, 128 rows}; nothing is transposed or re-laid-out:
//   S = Q K^T : A = Q tile (K-major, 128B swizzle), B = K rows (K-major)  -> fp32 S in TMEM, one column per key
//   P V       : A = P from TENSOR MEMORY (16-bit, two keys per column, written by tcgen05.st over the S columns
//               it replaces), B = V rows as they lie in memory = MN-major operand (head_dim contiguous)
// Because a whole segment's keys fit in TMEM (<= 512 columns) the softmax is exact two-pass (row max, then
// exp / sum) with no online rescaling.  One CTA works on one (segment, head) at a time: K and V are loaded once
// and reused by all of its 128-row query tiles.
//
// Roles (256 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4-7 softmax + output (thread <-> query row = TMEM lane).
#include "common.cuh"
#include "host_common.h"

#include <math.h>

namespace attn {

constexpr int D = 64;            // head dim
constexpr int BLOCK_Q = 128;
constexpr int MAX_LK = 512;
constexpr int THREADS = 256;
constexpr int Q_BYTES = BLOCK_Q * D * 2;        // 16 KB
constexpr int KV_BYTES = MAX_LK * D * 2;        // 64 KB each
constexpr int SMEM_BYTES = Q_BYTES + 2 * KV_BYTES + 1024;
constexpr int TMEM_COLS = 512;
constexpr float LOG2E = 1.4426950408889634f;

struct Params {
    int B, H, Lq, Lk;
    int q_col0, k_col0, v_col0;  // column of head 0 in the Q / K / V row buffers
    uint16_t* O;
    int64_t ldo;
    const float* add_mask;    // [B, Lk] additive key mask (0 or -10000 / -1e9 ...) or nullptr
    const float* bias_delta;  // [H, Lq + Lk - 1]: bias for (j - i) + (Lq - 1), or nullptr
    float scale;
    float causal_value;       // 0 = not causal; otherwise the additive value for j > i (reference: -10000)
    // split-KV (decoder cross-attention over n_ctx*L keys): segment b reads the queries of batch b / q_div and
    // writes UN-normalised fp32 partial outputs + (row max, row sum) for combine_splits_kernel
    int q_div;
    float* o_partial;         // [B*Lq, H*64] fp32 or nullptr
    float* ml_partial;        // [B*Lq, H, 2] fp32
};

__device__ __forceinline__ void tmem_ld32f(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    ab::tmem_ld32(taddr, r);
    ab::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

// smem descriptor of an MN-major operand tile stored as rows of 128 bytes (64 x 16-bit along MN) with the 128B
// swizzle: 8-row (K) groups are 1024 bytes apart (SBO); LBO (stride between 64-element MN blocks) is unused for N = 64.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1024 >> 4) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

template <bool kBF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    if constexpr (kBF16) {
        return static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16_rn(a))) |
               (static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16_rn(b))) << 16);
    }
    return static_cast<uint32_t>(__half_as_ushort(__float2half_rn(a))) |
           (static_cast<uint32_t>(__half_as_ushort(__float2half_rn(b))) << 16);
}

template <bool kBF16>
__global__ void __launch_bounds__(THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t kv_full, kv_empty, q_full, q_empty, s_full, p_ready, o_full, s_free;
    __shared__ uint32_t tmem_base_smem;
    __shared__ float s_bias[2 * MAX_LK];  // bias by (j - i) + (Lq - 1), this head
    __shared__ float s_mask[MAX_LK];      // additive key mask, this segment

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t smem_base = (ab::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - ab::smem_u32(smem_raw));
    uint8_t* sQ = smem_gen;
    uint8_t* sK = smem_gen + Q_BYTES;
    uint8_t* sV = sK + KV_BYTES;
    const uint32_t aQ = smem_base, aK = smem_base + Q_BYTES, aV = aK + KV_BYTES;

    const int n_chunks = (p.Lk + 127) / 128;   // 128-key chunks (TMA boxes / S column blocks)
    const int lk_pad = n_chunks * 128;
    const int n_qt = (p.Lq + BLOCK_Q - 1) / BLOCK_Q;
    const int n_items = p.B * p.H;
    const uint32_t o_col = static_cast<uint32_t>(lk_pad / 2);  // O accumulator: right after the packed P columns

    if (warp == 0 && lane == 0) {
        ab::tma_prefetch_desc(&tmap_q);
        ab::tma_prefetch_desc(&tmap_k);
        ab::tma_prefetch_desc(&tmap_v);
    }
    if (warp == 1 && lane == 0) {
        ab::mbar_init(&kv_full, 1);
        ab::mbar_init(&kv_empty, 1);
        ab::mbar_init(&q_full, 1);
        ab::mbar_init(&q_empty, 1);
        ab::mbar_init(&s_full, 1);
        ab::mbar_init(&p_ready, 128);
        ab::mbar_init(&o_full, 1);
        ab::mbar_init(&s_free, 128);
        ab::fence_barrier_init();
    }
    if (warp == 2) ab::tmem_alloc<1>(&tmem_base_smem, TMEM_COLS);
    ab::tc_fence_before();
    __syncthreads();
    ab::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int item_it = 0, qt_it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_it) {
                const int b = item / p.H, h = item % p.H;
                ab::mbar_wait(&kv_empty, (item_it & 1) ^ 1u, 21);
                ab::mbar_arrive_expect_tx(&kv_full, static_cast<uint32_t>(2 * n_chunks * 128 * D * 2));
                for (int c = 0; c < n_chunks; ++c) {
                    ab::tma_load_2d(&tmap_k, &kv_full, sK + c * (128 * D * 2), p.k_col0 + h * D, b * p.Lk + c * 128,
                                    ab::kEvictNormal);
                    ab::tma_load_2d(&tmap_v, &kv_full, sV + c * (128 * D * 2), p.v_col0 + h * D, b * p.Lk + c * 128,
                                    ab::kEvictNormal);
                }
                for (int qt = 0; qt < n_qt; ++qt, ++qt_it) {
                    ab::mbar_wait(&q_empty, (qt_it & 1) ^ 1u, 22);
                    ab::mbar_arrive_expect_tx(&q_full, Q_BYTES);
                    ab::tma_load_2d(&tmap_q, &q_full, sQ, p.q_col0 + h * D, (b / p.q_div) * p.Lq + qt * BLOCK_Q,
                                    ab::kEvictFirst);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_s = ab::umma_idesc_f16(BLOCK_Q, 128, kBF16);
            // P.V: A from TMEM (K-major), B = V rows = MN-major operand -> b_major bit (16) set
            constexpr uint32_t idesc_o = ab::umma_idesc_f16(BLOCK_Q, D, kBF16) | (1u << 16);
            int item_it = 0, qt_it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_it) {
                ab::mbar_wait(&kv_full, item_it & 1, 23);
                for (int qt = 0; qt < n_qt; ++qt, ++qt_it) {
                    ab::mbar_wait(&q_full, qt_it & 1, 24);
                    ab::mbar_wait(&s_free, (qt_it & 1) ^ 1u, 25);
                    ab::tc_fence_after();
                    const uint64_t qdesc = ab::umma_desc_k_sw128(aQ);
                    for (int c = 0; c < n_chunks; ++c) {
                        const uint64_t kdesc = ab::umma_desc_k_sw128(aK + c * (128 * D * 2));
#pragma unroll
                        for (int k = 0; k < D / 16; ++k)
                            ab::umma_ss<1>(tmem_base + c * 128, qdesc + ((k * 32) >> 4), kdesc + ((k * 32) >> 4), idesc_s,
                                           k != 0 ? 1u : 0u);
                    }
                    ab::umma_commit(&q_empty);  // Q tile consumed once the S MMAs retire
                    ab::umma_commit(&s_full);
                    ab::mbar_wait(&p_ready, qt_it & 1, 26);
                    ab::tc_fence_after();
                    const uint64_t vdesc = umma_desc_mn_sw128(aV);
                    for (int k = 0; k < lk_pad / 16; ++k)
                        ab::umma_ts<1>(tmem_base + o_col, tmem_base + k * 8, vdesc + static_cast<uint64_t>((k * 2048) >> 4),
                                       idesc_o, k != 0 ? 1u : 0u);
                    if (qt == n_qt - 1) ab::umma_commit(&kv_empty);  // last use of this (segment, head)'s K / V
                    ab::umma_commit(&o_full);
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== softmax + output =====================
        const uint32_t lg = warp & 3u;
        const int row_in_tile = static_cast<int>(lg * 32 + lane);
        const uint32_t lane_addr = tmem_base + ((lg * 32u) << 16);
        const int tid = static_cast<int>(threadIdx.x) - 128;
        int qt_it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int b = item / p.H, h = item % p.H;
            // per-(segment, head) tables; the previous item's softmax is finished (all 128 threads passed its o_full)
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int j = tid; j < lk_pad; j += 128)
                s_mask[j] = (j < p.Lk) ? (p.add_mask ? p.add_mask[static_cast<size_t>(b) * p.Lk + j] : 0.f) : -INFINITY;
            if (p.bias_delta)
                for (int d = tid; d < p.Lq + p.Lk - 1; d += 128)
                    s_bias[d] = p.bias_delta[static_cast<size_t>(h) * (p.Lq + p.Lk - 1) + d];
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int qt = 0; qt < n_qt; ++qt, ++qt_it) {
                const int i = qt * BLOCK_Q + row_in_tile;       // query position inside the segment
                const int boff = p.Lq - 1 - i;                  // bias index = j + boff
                ab::mbar_wait(&s_full, qt_it & 1, 27);
                ab::tc_fence_after();
                // ---- pass 1: row max of scale*s + bias + mask ----
                float mx = -INFINITY;
                for (int c = 0; c < lk_pad / 32; ++c) {
                    float v[32];
                    tmem_ld32f(lane_addr + c * 32, v);
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) {
                        const int j = c * 32 + jj;
                        float s = v[jj] * p.scale + s_mask[j];
                        if (p.bias_delta) s += s_bias[min(max(j + boff, 0), 2 * MAX_LK - 1)];
                        if (p.causal_value != 0.f && j > i) s += p.causal_value;
                        mx = fmaxf(mx, s);
                    }
                }
                // ---- pass 2: p = exp(s - max), row sum, P (16-bit) written over the S columns it replaces ----
                float sum = 0.f;
                const float mxl = mx * LOG2E;
                for (int c = 0; c < lk_pad / 32; ++c) {
                    float v[32];
                    tmem_ld32f(lane_addr + c * 32, v);
                    uint32_t pk[16];
#pragma unroll
                    for (int jj = 0; jj < 32; jj += 2) {
                        float e[2];
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int j = c * 32 + jj + u;
                            float s = v[jj + u] * p.scale + s_mask[j];
                            if (p.bias_delta) s += s_bias[min(max(j + boff, 0), 2 * MAX_LK - 1)];
                            if (p.causal_value != 0.f && j > i) s += p.causal_value;
                            e[u] = exp2f(s * LOG2E - mxl);
                            sum += e[u];
                        }
                        pk[jj >> 1] = pack2<kBF16>(e[0], e[1]);
                    }
                    tmem_st16(lane_addr + c * 16, pk);
                }
                ab::tmem_st_wait();
                ab::tc_fence_before();
                ab::mbar_arrive(&p_ready);
                // ---- output: O / sum -> 16-bit, 128 contiguous bytes per row ----
                ab::mbar_wait(&o_full, qt_it & 1, 28);
                ab::tc_fence_after();
                if (p.o_partial != nullptr) {
                    // split-KV: un-normalised partial output in fp32 + (max, sum) of this split
                    float* dst = p.o_partial + (static_cast<size_t>(b) * p.Lq + i) * (static_cast<size_t>(p.H) * D) + h * D;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        float v[32];
                        tmem_ld32f(lane_addr + o_col + c * 32, v);
                        if (i < p.Lq) {
#pragma unroll
                            for (int v4 = 0; v4 < 8; ++v4)
                                reinterpret_cast<float4*>(dst + c * 32)[v4] =
                                    make_float4(v[4 * v4], v[4 * v4 + 1], v[4 * v4 + 2], v[4 * v4 + 3]);
                        }
                    }
                    ab::tc_fence_before();
                    ab::mbar_arrive(&s_free);
                    if (i < p.Lq) {
                        float* ml = p.ml_partial + ((static_cast<size_t>(b) * p.Lq + i) * p.H + h) * 2;
                        ml[0] = mx;
                        ml[1] = sum;
                    }
                } else {
                const float inv = 1.0f / sum;
                uint32_t outw[32];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    float v[32];
                    tmem_ld32f(lane_addr + o_col + c * 32, v);
#pragma unroll
                    for (int jj = 0; jj < 32; jj += 2) outw[c * 16 + (jj >> 1)] = pack2<kBF16>(v[jj] * inv, v[jj + 1] * inv);
                }
                ab::tc_fence_before();
                ab::mbar_arrive(&s_free);  // S / P / O columns may be overwritten by the next query tile
                if (i < p.Lq) {
                    uint4* dst = reinterpret_cast<uint4*>(p.O + (static_cast<size_t>(b) * p.Lq + i) * p.ldo + h * D);
#pragma unroll
                    for (int v4 = 0; v4 < 8; ++v4)
                        dst[v4] = make_uint4(outw[4 * v4], outw[4 * v4 + 1], outw[4 * v4 + 2], outw[4 * v4 + 3]);
                }
                }
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

// out[b, i, h, :] = sum_s w_s O_s / sum_s w_s l_s with w_s = exp(m_s - max_s m_s); one warp per (b, i, h)
template <bool kBF16>
__global__ void combine_splits_kernel(const float* __restrict__ o_partial, const float* __restrict__ ml, int B, int splits,
                                      int Lq, int H, uint16_t* __restrict__ out, int64_t ldo) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= B * Lq * H) return;
    const int h = w % H, i = (w / H) % Lq, b = w / (H * Lq);
    float M = -INFINITY;
    for (int s = 0; s < splits; ++s) {
        const size_t row = (static_cast<size_t>(b) * splits + s) * Lq + i;
        M = fmaxf(M, ml[(row * H + h) * 2]);
    }
    float acc0 = 0.f, acc1 = 0.f, den = 0.f;
    for (int s = 0; s < splits; ++s) {
        const size_t row = (static_cast<size_t>(b) * splits + s) * Lq + i;
        const float wgt = exp2f((ml[(row * H + h) * 2] - M) * LOG2E);
        den += wgt * ml[(row * H + h) * 2 + 1];
        const float2 o = reinterpret_cast<const float2*>(o_partial + row * (static_cast<size_t>(H) * D) + h * D)[lane];
        acc0 += wgt * o.x;
        acc1 += wgt * o.y;
    }
    const float inv = 1.0f / den;
    reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(b) * Lq + i) * ldo + h * D)[lane] =
        pack2<kBF16>(acc0 * inv, acc1 * inv);
}

}  // namespace attn

extern "C" {

int atlas_b200_attention(const void* q, int64_t ldq, int32_t q_col0, const void* k, int64_t ldk, int32_t k_col0,
                         const void* v, int64_t ldv, int32_t v_col0, void* out, int64_t ldo, const float* add_mask,
                         const float* bias_delta, int32_t B, int32_t H, int32_t Lq, int32_t Lk, float scale,
                         float causal_value, int32_t q_div, float* o_partial, float* ml_partial, int32_t is_bf16,
                         void* stream) {
    using namespace attn;
    AB_REQUIRE(q_div >= 1 && B % q_div == 0, "attention: q_div must divide the number of key segments");
    AB_REQUIRE((o_partial == nullptr) == (ml_partial == nullptr), "attention: partial outputs come in pairs");
    AB_REQUIRE(B >= 0 && H > 0 && Lq > 0 && Lk > 0 && Lk <= MAX_LK, "attention: need 0 < Lk <= %d (got Lq=%d Lk=%d)",
               MAX_LK, Lq, Lk);
    AB_REQUIRE(Lq + Lk - 1 <= 2 * MAX_LK, "attention: Lq + Lk too large for the bias table");
    AB_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && q_col0 % 8 == 0 && k_col0 % 8 == 0 &&
                   v_col0 % 8 == 0,
               "attention: strides and column offsets must be multiples of 8 elements");
    if (B == 0) return ATLAS_B200_OK;
    CUtensorMap tq, tk, tv;
    int rc = abh::make_tmap_2d_16bit(&tq, q, static_cast<uint64_t>(B / q_div) * Lq, static_cast<uint64_t>(q_col0 + H * D),
                                     static_cast<uint64_t>(ldq), BLOCK_Q, D, is_bf16 != 0);
    if (rc) return rc;
    rc = abh::make_tmap_2d_16bit(&tk, k, static_cast<uint64_t>(B) * Lk, static_cast<uint64_t>(k_col0 + H * D),
                                 static_cast<uint64_t>(ldk), 128, D, is_bf16 != 0);
    if (rc) return rc;
    rc = abh::make_tmap_2d_16bit(&tv, v, static_cast<uint64_t>(B) * Lk, static_cast<uint64_t>(v_col0 + H * D),
                                 static_cast<uint64_t>(ldv), 128, D, is_bf16 != 0);
    if (rc) return rc;
    Params p;
    p.B = B;
    p.H = H;
    p.Lq = Lq;
    p.Lk = Lk;
    p.q_col0 = q_col0;
    p.k_col0 = k_col0;
    p.v_col0 = v_col0;
    p.O = static_cast<uint16_t*>(out);
    p.ldo = ldo;
    p.add_mask = add_mask;
    p.bias_delta = bias_delta;
    p.scale = scale;
    p.causal_value = causal_value;
    p.q_div = q_div;
    p.o_partial = o_partial;
    p.ml_partial = ml_partial;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    static bool attr_set[2] = {false, false};
    const int items = B * H;
    const int grid = items < abh::num_sms() ? items : abh::num_sms();
    if (is_bf16) {
        if (!attr_set[1]) {
            AB_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               SMEM_BYTES));
            attr_set[1] = true;
        }
        attention_kernel<true><<<grid, THREADS, SMEM_BYTES, s>>>(tq, tk, tv, p);
    } else {
        if (!attr_set[0]) {
            AB_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               SMEM_BYTES));
            attr_set[0] = true;
        }
        attention_kernel<false><<<grid, THREADS, SMEM_BYTES, s>>>(tq, tk, tv, p);
    }
    abh::count_launch();
    AB_CUDA_CHECK(cudaGetLastError());
    return ATLAS_B200_OK;
}

int atlas_b200_attention_combine(const float* o_partial, const float* ml_partial, int32_t B, int32_t splits, int32_t Lq,
                                 int32_t H, void* out, int64_t ldo, int32_t is_bf16, void* stream) {
    AB_REQUIRE(B >= 0 && splits >= 1 && Lq > 0 && H > 0 && ldo % 2 == 0, "attention_combine: bad shape");
    if (B == 0) return ATLAS_B200_OK;
    const int warps = B * Lq * H;
    const int grid = (warps * 32 + 255) / 256;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (is_bf16)
        attn::combine_splits_kernel<true><<<grid, 256, 0, s>>>(o_partial, ml_partial, B, splits, Lq, H,
                                                               static_cast<uint16_t*>(out), ldo);
    else
        attn::combine_splits_kernel<false><<<grid, 256, 0, s>>>(o_partial, ml_partial, B, splits, Lq, H,
                                                                static_cast<uint16_t*>(out), ldo);
    abh::count_launch();
    AB_CUDA_CHECK(cudaGetLastError());
    return ATLAS_B200_OK;
}

}  // extern "C"
