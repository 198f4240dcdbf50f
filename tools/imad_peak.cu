// Integer-multiply roofline microbenchmark for B200 ("No integer peak is recorded -- measure it").
// Every kernel changes a multiplicand each iteration so ptxas cannot hoist the product (an earlier version with
// loop-invariant operands was optimised into 64-bit adds and reported a bogus 2x figure).
//   wide_cols   mad.wide.u32 with 64-bit addend, 13 independent column accumulators (IMAD.WIDE.U32)
//   wide_chain  carry chains of mad.lo.cc/madc.hi.cc pairs, K chains of 6 (IMAD.WIDE.U32.X) -- the Montgomery row pattern
//   imad32      mad.lo.u32 (IMAD)         imadhi   mad.hi.u32 (IMAD.HI.U32)
//   fpmul       FpOps<C>::mul dependent chains, many threads (the kernel building block)
// Prints one JSON object; "imad_wide_realistic_per_s" is the roofline denominator used by bench.py / DESIGN.md.
#include <cstdio>
#include <cuda_runtime.h>
#include "../mathlib_b200/csrc/curves.cuh"
using namespace b200;

template <int K>
__global__ void k_chain(uint32_t* out, uint32_t a, uint32_t b, int iters) {
    uint32_t acc[K][12];
    uint32_t x[6];
#pragma unroll
    for (int k = 0; k < K; k++)
#pragma unroll
        for (int j = 0; j < 12; j++) acc[k][j] = threadIdx.x + j + k;
#pragma unroll
    for (int j = 0; j < 6; j++) x[j] = a + threadIdx.x * (j + 1);
    uint32_t y = b + blockIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            asm volatile(
                "mad.lo.cc.u32 %0, %12, %18, %0;\n\t madc.hi.cc.u32 %1, %12, %18, %1;\n\t"
                "madc.lo.cc.u32 %2, %13, %18, %2;\n\t madc.hi.cc.u32 %3, %13, %18, %3;\n\t"
                "madc.lo.cc.u32 %4, %14, %18, %4;\n\t madc.hi.cc.u32 %5, %14, %18, %5;\n\t"
                "madc.lo.cc.u32 %6, %15, %18, %6;\n\t madc.hi.cc.u32 %7, %15, %18, %7;\n\t"
                "madc.lo.cc.u32 %8, %16, %18, %8;\n\t madc.hi.cc.u32 %9, %16, %18, %9;\n\t"
                "madc.lo.cc.u32 %10, %17, %18, %10;\n\t madc.hi.u32 %11, %17, %18, %11;\n\t"
                : "+r"(acc[k][0]), "+r"(acc[k][1]), "+r"(acc[k][2]), "+r"(acc[k][3]), "+r"(acc[k][4]), "+r"(acc[k][5]),
                  "+r"(acc[k][6]), "+r"(acc[k][7]), "+r"(acc[k][8]), "+r"(acc[k][9]), "+r"(acc[k][10]), "+r"(acc[k][11])
                : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(y));
        }
        y += 0x9e3779b9u;
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < K; k++)
#pragma unroll
        for (int j = 0; j < 12; j++) s ^= acc[k][j];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int K>
__global__ void k_cols(uint64_t* out, uint32_t a, uint32_t b, int iters) {
    uint64_t acc[K];
    uint32_t x[K];
#pragma unroll
    for (int k = 0; k < K; k++) { acc[k] = threadIdx.x + k; x[k] = a + threadIdx.x * (k + 1); }
    uint32_t y = b + blockIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < K; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(x[k]), "r"(y));
        y += 0x9e3779b9u;
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < K; k++) s ^= acc[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int K, bool HI>
__global__ void k_imad32(uint32_t* out, uint32_t a, uint32_t b, int iters) {
    uint32_t acc[K], x[K];
#pragma unroll
    for (int k = 0; k < K; k++) { acc[k] = threadIdx.x + k; x[k] = a + threadIdx.x * (k + 1); }
    uint32_t y = b + blockIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            if (HI) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(x[k]), "r"(y));
            else asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(x[k]), "r"(y));
        }
        y += 0x9e3779b9u;
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < K; k++) s ^= acc[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class C>
__global__ void k_fpmul(Fp<C::N>* io, int iters) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    Fp<C::N> x = io[t], y = io[t];
    y.l[0] ^= 1;
    for (int i = 0; i < iters; i++) { FpOps<C>::mul(x, x, y); }
    io[t] = x;
}

static float tms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }
template <class F> static double rate(F launch, double ops) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    double best = 0;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        double v = ops / (tms(e0, e1) * 1e-3);
        if (v > best) best = v;
    }
    return best;
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int sms = pr.multiProcessorCount;
    void* buf; cudaMalloc(&buf, (size_t)sms * 2048 * 64);
    cudaMemset(buf, 1, (size_t)sms * 2048 * 64);
    const int iters = 20000;
    double best_cols = 0, best_chain = 0, best_lo = 0, best_hi = 0;
    for (int tps : {256, 512, 1024}) {
        int tpb = 128, blocks = sms * (tps / tpb);
        double n = (double)blocks * tpb * iters;
        best_cols = fmax(best_cols, rate([&] { k_cols<13><<<blocks, tpb>>>((uint64_t*)buf, 3, 5, iters); }, n * 13));
        best_chain = fmax(best_chain, rate([&] { k_chain<4><<<blocks, tpb>>>((uint32_t*)buf, 3, 5, iters); }, n * 24));
        best_lo = fmax(best_lo, rate([&] { k_imad32<8, false><<<blocks, tpb>>>((uint32_t*)buf, 3, 5, iters); }, n * 8));
        best_hi = fmax(best_hi, rate([&] { k_imad32<8, true><<<blocks, tpb>>>((uint32_t*)buf, 3, 5, iters); }, n * 8));
    }
    double nominal = 32.0 * sms * pr.clockRate * 1e3;   // fmaheavy: one IMAD.WIDE warp-instruction per 4 cycles per SMSP
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"imad_wide_nominal_per_s\": %.4e, \"imad_wide_cols_per_s\": %.4e, "
           "\"imad_wide_chain_per_s\": %.4e, \"imad_wide_realistic_per_s\": %.4e, \"imad_lo_per_s\": %.4e, \"imad_hi_per_s\": %.4e",
           pr.name, sms, pr.clockRate, nominal, best_cols, best_chain, fmax(best_cols, best_chain), best_lo, best_hi);
    printf(", \"fpmul\": [");
    bool first = true;
    for (int tps : {128, 256, 512, 1024}) {
        int tpb = 128, blocks = sms * (tps / tpb);
        const int it2 = 2000;
        double m12 = rate([&] { k_fpmul<BLS381><<<blocks, tpb>>>((Fp<12>*)buf, it2); }, (double)blocks * tpb * it2);
        double m8 = rate([&] { k_fpmul<BN254><<<blocks, tpb>>>((Fp<8>*)buf, it2); }, (double)blocks * tpb * it2);
        printf("%s{\"threads_per_sm\": %d, \"bls381_mul_per_s\": %.4e, \"bn254_mul_per_s\": %.4e}", first ? "" : ", ", tps, m12, m8);
        first = false;
    }
    printf("]}\n");
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}
