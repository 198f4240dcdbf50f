// Integer-multiply roofline microbenchmark for B200 ("No integer peak is recorded -- measure it").
// Measures (a) dependency-free IMAD.WIDE.U32 issue rate, (b) 32-bit IMAD rate, (c) Montgomery Fp-mul throughput
// of fp.cuh at several occupancies.  Prints one JSON object; bench.py / DESIGN.md use "imad_wide_per_s".
#include <cstdio>
#include <cuda_runtime.h>
#include "../mathlib_b200/csrc/curves.cuh"
using namespace b200;

template <int ILP>
__global__ void k_imad_wide(uint64_t* out, uint32_t a, uint32_t b, int iters) {
    uint64_t acc[ILP];
    uint32_t x = a + threadIdx.x, y = b + blockIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = i + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(x), "r"(y));
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s ^= acc[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_imad32(uint32_t* out, uint32_t a, uint32_t b, int iters) {
    uint32_t acc[ILP];
    uint32_t x = a + threadIdx.x, y = b + blockIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = i + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(x), "r"(y));
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s ^= acc[i];
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

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int sms = pr.multiProcessorCount;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    void* buf; cudaMalloc(&buf, (size_t)sms * 32 * 1024 * 64);
    cudaMemset(buf, 1, (size_t)sms * 32 * 1024 * 64);
    printf("{\"gpu\": \"%s\", \"sms\": %d", pr.name, sms);
    const int iters = 4096;
    double best_wide = 0, best_32 = 0;
    for (int tpb : {256, 512, 1024}) {
        int blocks = sms * (2048 / tpb);
        k_imad_wide<8><<<blocks, tpb>>>((uint64_t*)buf, 3, 5, 16);
        cudaEventRecord(e0);
        k_imad_wide<8><<<blocks, tpb>>>((uint64_t*)buf, 3, 5, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        double ops = (double)blocks * tpb * iters * 8 * 8;
        double r = ops / (time_ms(e0, e1) * 1e-3);
        if (r > best_wide) best_wide = r;
        k_imad32<8><<<blocks, tpb>>>((uint32_t*)buf, 3, 5, 16);
        cudaEventRecord(e0);
        k_imad32<8><<<blocks, tpb>>>((uint32_t*)buf, 3, 5, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        r = ops / (time_ms(e0, e1) * 1e-3);
        if (r > best_32) best_32 = r;
    }
    printf(", \"imad_wide_per_s\": %.4e, \"imad32_per_s\": %.4e", best_wide, best_32);
    // sustained (2 s) IMAD.WIDE rate
    {
        int tpb = 512, blocks = sms * 4;
        cudaEventRecord(e0);
        int reps = 0;
        for (; reps < 400; reps++) k_imad_wide<8><<<blocks, tpb>>>((uint64_t*)buf, 3, 5, iters * 4);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        double ops = (double)reps * blocks * tpb * iters * 4 * 64;
        printf(", \"imad_wide_sustained_per_s\": %.4e, \"sustained_ms\": %.1f", ops / (time_ms(e0, e1) * 1e-3), time_ms(e0, e1));
    }
    // Fp mul throughput
    printf(", \"fpmul\": [");
    bool first = true;
    for (int tpb : {64, 128, 256, 512}) {
        for (int bps : {1, 2, 4, 8}) {
            if (tpb * bps > 2048) continue;
            int blocks = sms * bps;
            const int it2 = 2000;
            k_fpmul<BLS381><<<blocks, tpb>>>((Fp<12>*)buf, 10);
            cudaEventRecord(e0);
            k_fpmul<BLS381><<<blocks, tpb>>>((Fp<12>*)buf, it2);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            double m12 = (double)blocks * tpb * it2 / (time_ms(e0, e1) * 1e-3);
            k_fpmul<BN254><<<blocks, tpb>>>((Fp<8>*)buf, 10);
            cudaEventRecord(e0);
            k_fpmul<BN254><<<blocks, tpb>>>((Fp<8>*)buf, it2);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            double m8 = (double)blocks * tpb * it2 / (time_ms(e0, e1) * 1e-3);
            printf("%s{\"threads_per_sm\": %d, \"tpb\": %d, \"bls381_mul_per_s\": %.4e, \"bn254_mul_per_s\": %.4e}", first ? "" : ", ",
                   tpb * bps, tpb, m12, m8);
            first = false;
        }
    }
    printf("]}\n");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "cuda error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
