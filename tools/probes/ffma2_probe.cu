// Microbenchmark (B200): issue cost of packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2, sm_100a) against scalar FFMA, alone and
// interleaved with MUFU -- what the instruction-issue-bound march kernels (K6, K16) would gain from packing two steps / two
// channels into one instruction.  Build + run: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_probe tools/probes/ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { return (u64)__float_as_uint(a) | ((u64)__float_as_uint(b) << 32); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
constexpr int N = 4096;
template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, float s) {
    float a[8]; u64 p[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3f + i; p[i] = pk(a[i], a[i] + 0.5f); }
    float m = 1.0f + threadIdx.x * 1e-6f;
    const u64 ss = pk(s, s), cc = pk(0.25f, 0.125f);
#pragma unroll 1
    for (int it = 0; it < N; ++it) {
        if (MODE == 0) {          // 16 scalar FFMA
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = fmaf(a[i], s, 0.25f); }
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = fmaf(a[i], s, 0.125f); }
        } else if (MODE == 1) {   // 8 FFMA2 = the same 16 FMAs
#pragma unroll
            for (int i = 0; i < 8; ++i) { p[i] = fma2(p[i], ss, cc); }
        } else if (MODE == 2) {   // 16 scalar FFMA + 2 MUFU
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = fmaf(a[i], s, 0.25f); }
            m = __frcp_rn(m) ; m = __fsqrt_rn(m);
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = fmaf(a[i], s, 0.125f); }
        } else if (MODE == 3) {   // 8 FFMA2 + 2 MUFU
#pragma unroll
            for (int i = 0; i < 8; ++i) { p[i] = fma2(p[i], ss, cc); }
            m = __frcp_rn(m); m = __fsqrt_rn(m);
        } else if (MODE == 4) {   // 16 FFMA2 = 32 FMAs
#pragma unroll
            for (int i = 0; i < 8; ++i) { p[i] = fma2(p[i], ss, cc); }
#pragma unroll
            for (int i = 0; i < 8; ++i) { p[i] = fma2(p[i], ss, cc); }
        } else if (MODE == 6) {   // 16 scalar FFMA + 16 integer ALU ops
            unsigned q = __float_as_uint(m);
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = fmaf(a[i], s, 0.25f); q = (q ^ (q >> 3)) + i; }
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = fmaf(a[i], s, 0.125f); }
            m = __uint_as_float(q);
        } else if (MODE == 7) {   // 8 FFMA2 + the same 16 integer ALU ops
            unsigned q = __float_as_uint(m);
#pragma unroll
            for (int i = 0; i < 8; ++i) { p[i] = fma2(p[i], ss, cc); q = (q ^ (q >> 3)) + i; }
            m = __uint_as_float(q);
        } else if (MODE == 8) {   // 16 scalar FFMA + 2 approximate MUFU
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = fmaf(a[i], s, 0.25f); }
            m = __fdividef(1.0f, m); a[7] = rsqrtf(a[7]);
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = fmaf(a[i], s, 0.125f); }
        } else if (MODE == 9) {   // 8 FFMA2 + 2 approximate MUFU
#pragma unroll
            for (int i = 0; i < 8; ++i) { p[i] = fma2(p[i], ss, cc); }
            m = __fdividef(1.0f, m); a[7] = rsqrtf(a[7]);
        } else if (MODE == 10) {  // 8 packed mul + 8 packed add
#pragma unroll
            for (int i = 0; i < 8; ++i) { asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ss)); }
#pragma unroll
            for (int i = 0; i < 8; ++i) { asm("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc)); }
        } else if (MODE == 5) {   // 2 approximate MUFU (rcp + rsqrt), independent chains
            m = __fdividef(1.0f, m) + 1.0f; a[0] = rsqrtf(a[0]) + 1.0f; a[1] = __fdividef(1.0f, a[1]) + 1.0f; a[2] = rsqrtf(a[2]) + 1.0f;
        }
    }
    float r = m;
    for (int i = 0; i < 8; ++i) r += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run(const char* name, double fma_per_iter, float* out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8;
    probe<MODE><<<blocks, 256>>>(out, 0.999f);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) probe<MODE><<<blocks, 256>>>(out, 0.999f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    double warps = double(blocks) * 8, iters = double(N);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double cycles = ms * 1e-3 * clk * 1e3;                         // SM cycles of the launch (at the nominal max clock)
    double per_smsp = warps * iters / (148.0 * 4.0);                // loop iterations each SM sub-partition executes
    printf("%-28s %8.3f ms   %6.2f cycles / warp-iteration / SMSP   %7.1f TFMA/s\n", name, ms, cycles / per_smsp, fma_per_iter * warps * 32 * iters / (ms * 1e-3) * 1e-12);
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<0>("16 FFMA", 16, out);
    run<1>("8 FFMA2 (16 FMA)", 16, out);
    run<4>("16 FFMA2 (32 FMA)", 32, out);
    run<2>("16 FFMA + 2 MUFU", 16, out);
    run<3>("8 FFMA2 + 2 MUFU", 16, out);
    run<5>("4 MUFU (approx)", 0, out);
    run<6>("16 FFMA + 16 int ALU", 16, out);
    run<7>("8 FFMA2 + 16 int ALU", 16, out);
    run<8>("16 FFMA + 2 MUFU approx", 16, out);
    run<9>("8 FFMA2 + 2 MUFU approx", 16, out);
    run<10>("8 FMUL2 + 8 FADD2", 16, out);
    return 0;
}
