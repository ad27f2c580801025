// Micro-benchmark of the softmax inner loops of csrc/attention_lanes.cu WITHOUT tensor cores / TMEM / barriers:
// how many cycles per score element does an SM need as a function of resident warps and of the instruction mix?
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o softmax_ubench softmax_ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int BK = 48;
__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
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
__device__ __forceinline__ uint32_t pack2(float a, float b) { uint32_t d; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a)); return d; }

// MODE bits: 1 = bias from shared memory (LDS.128), 2 = MUFU exp, 4 = packed f32x2 arithmetic (else scalar), 8 = pack to bf16
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float* out, const float* in, int iters) {
    __shared__ __align__(16) float s_bias[4 * 1160];
    for (int i = threadIdx.x; i < 4 * 1160; i += blockDim.x) s_bias[i] = 0.001f * (i & 63);
    __syncthreads();
    const int row = threadIdx.x & 127;
    const int off = 383 - row;
    const float* pb_row = s_bias + (off & 3) * 1160 + (off & ~3);
    float r[BK];
    float base = in[threadIdx.x & 31];
    float m_ref = 0.f, sum[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int e = 0; e < BK; ++e) r[e] = base + 0.01f * e + 1e-3f * it;
        const float* pb = pb_row + (it & 7) * BK;
        float4 add[BK / 4];
#pragma unroll
        for (int q = 0; q < BK / 4; ++q) {
            if (MODE & 1) add[q] = *reinterpret_cast<const float4*>(pb + 4 * q);
            else add[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float mx[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
        for (int q = 0; q < BK / 4; ++q) {
            float t0, t1, t2, t3;
            if (MODE & 4) {
                ffma2(t0, t1, r[4 * q], r[4 * q + 1], 1.44f, add[q].x, add[q].y);
                ffma2(t2, t3, r[4 * q + 2], r[4 * q + 3], 1.44f, add[q].z, add[q].w);
            } else {
                t0 = fmaf(r[4 * q], 1.44f, add[q].x); t1 = fmaf(r[4 * q + 1], 1.44f, add[q].y);
                t2 = fmaf(r[4 * q + 2], 1.44f, add[q].z); t3 = fmaf(r[4 * q + 3], 1.44f, add[q].w);
            }
            r[4 * q] = t0; r[4 * q + 1] = t1; r[4 * q + 2] = t2; r[4 * q + 3] = t3;
            mx[q & 1] = fmax3(mx[q & 1], t0, t1);
            mx[2 + (q & 1)] = fmax3(mx[2 + (q & 1)], t2, t3);
        }
        const float mb = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        if (it == 0) m_ref = mb;
        const float neg_m = -m_ref;
#pragma unroll
        for (int jj = 0; jj < BK; jj += 2) {
            if (MODE & 4) fadd2(r[jj], r[jj + 1], r[jj], r[jj + 1], neg_m, neg_m);
            else { r[jj] += neg_m; r[jj + 1] += neg_m; }
        }
        if (MODE & 2) {
#pragma unroll
            for (int jj = 0; jj < BK; ++jj) r[jj] = ex2a(r[jj]);
        }
#pragma unroll
        for (int jj = 0; jj < BK; jj += 2) {
            const int a = (jj >> 1) & 1;
            if (MODE & 4) fadd2(sum[2 * a], sum[2 * a + 1], sum[2 * a], sum[2 * a + 1], r[jj], r[jj + 1]);
            else { sum[2 * a] += r[jj]; sum[2 * a + 1] += r[jj + 1]; }
            if (MODE & 8) acc ^= pack2(r[jj], r[jj + 1]);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum[0] + sum[1] + sum[2] + sum[3] + __uint_as_float(acc & 0xFF);
}

template <int MODE>
void run(const char* name, int warps, float* out, float* in) {
    const int iters = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, warps * 32>>>(out, in, 10);
    cudaEventRecord(e0);
    k<MODE><<<148, warps * 32>>>(out, in, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double elems_per_sm = double(warps) * 32 * BK * iters;
    printf("%-34s warps/SM %2d: %.3f ms  -> %.2f elements/ns/SM = %.1f elem/clk/SM @1.9GHz (MUFU peak 16)\n", name, warps, ms,
           elems_per_sm / (ms * 1e6), elems_per_sm / (ms * 1e6) / 1.9);
}

int main() {
    float *out, *in;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
    for (int warps : {4, 8, 12, 16, 24, 32}) {
        run<15>("bias+exp+f32x2+pack (kernel's mix)", warps, out, in);
    }
    for (int warps : {12, 24}) {
        run<14>("no bias", warps, out, in);
        run<13>("no exp", warps, out, in);
        run<11>("scalar fp32 (no f32x2)", warps, out, in);
        run<7>("no pack", warps, out, in);
        run<2>("exp only (+scalar adds)", warps, out, in);
        run<0>("adds / max only", warps, out, in);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
