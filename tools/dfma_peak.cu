// Gate experiment: could the FP64 pipe carry the multi-precision products instead of IMAD.WIDE?
// B200 issues DFMA at 64 lanes/clk/SM on its own pipe, IMAD.WIDE.U32 at 32 lanes/clk/SM.  Measured here:
//   dfma          independent fma.rn.f64 accumulators (raw pipe rate)
//   split52       the exact 52 x 52 -> 104-bit product primitive: hi = fma.rz(a, b, 2^104), lo = fma.rz(a, b, 2^104 - hi)
//                 (two DFMA + one DADD per limb product, operands varied every iteration)
//   mixed         split52 interleaved with IMAD.WIDE chains in the same warp: do the two pipes overlap?
// Prints one JSON object with instruction rates and bits^2/clk/SM for both multipliers (profiles/peaks_r2.json).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dfma_peak.bin tools/dfma_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int K>
__global__ void k_dfma(double* out, double a, double b, int iters) {
    double acc[K], x[K];
#pragma unroll
    for (int k = 0; k < K; k++) { acc[k] = threadIdx.x + k; x[k] = a + threadIdx.x * (k + 1); }
    double y = b + blockIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < K; k++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc[k]) : "d"(x[k]), "d"(y));
        y += 1.0;
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < K; k++) s += acc[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// K independent exact limb products per iteration, folded into K (hi, lo) accumulator pairs
template <int K>
__global__ void k_split52(double* out, double a, double b, int iters) {
    const double C = 20282409603651670423947251286016.0;      // 2^104
    double hi_acc[K], lo_acc[K], x[K];
#pragma unroll
    for (int k = 0; k < K; k++) { hi_acc[k] = 0; lo_acc[k] = 0; x[k] = a + threadIdx.x * (k + 1); }
    double y = b + blockIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            double hi, d, lo;
            asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(hi) : "d"(x[k]), "d"(y), "d"(C));
            asm volatile("sub.rz.f64 %0, %1, %2;" : "=d"(d) : "d"(C), "d"(hi));
            asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(lo) : "d"(x[k]), "d"(y), "d"(d));
            hi_acc[k] += hi;          // stand-in for the column accumulation (DADD)
            lo_acc[k] += lo;
        }
        y += 3.0;
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < K; k++) s += hi_acc[k] + lo_acc[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int K>
__global__ void k_wide(uint64_t* out, uint32_t a, uint32_t b, int iters) {
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
// both pipes from one instruction stream: KD DFMA accumulators and KW IMAD.WIDE accumulators per iteration
template <int KD, int KW>
__global__ void k_mixed(double* out, double a, double b, uint32_t ia, int iters) {
    double dacc[KD], dx[KD];
    uint64_t wacc[KW];
    uint32_t wx[KW];
#pragma unroll
    for (int k = 0; k < KD; k++) { dacc[k] = threadIdx.x + k; dx[k] = a + threadIdx.x * (k + 1); }
#pragma unroll
    for (int k = 0; k < KW; k++) { wacc[k] = threadIdx.x + k; wx[k] = ia + threadIdx.x * (k + 1); }
    double y = b + blockIdx.x;
    uint32_t iy = ia + blockIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < (KD > KW ? KD : KW); k++) {
            if (k < KD) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(dacc[k]) : "d"(dx[k]), "d"(y));
            if (k < KW) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(wacc[k]) : "r"(wx[k]), "r"(iy));
        }
        y += 1.0;
        iy += 0x9e3779b9u;
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < KD; k++) s += dacc[k];
#pragma unroll
    for (int k = 0; k < KW; k++) s += (double)wacc[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static double time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, threads = 256, blocks = sms * 8, iters = 4096;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    void* out;
    cudaMalloc(&out, (size_t)blocks * threads * 8);
    const double total = (double)blocks * threads * iters;
    double ms;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, khz);
    ms = time_ms([&] { k_dfma<16><<<blocks, threads>>>((double*)out, 1.5, 2.5, iters); });
    const double dfma = total * 16 / (ms * 1e-3);
    printf(", \"dfma_per_s\": %.4e, \"dfma_lanes_per_clk_per_sm\": %.1f", dfma, dfma / sms / (khz * 1e3));
    ms = time_ms([&] { k_split52<8><<<blocks, threads>>>((double*)out, 4503599627370495.0, 4503599627370001.0, iters); });
    const double split = total * 8 / (ms * 1e-3);
    printf(", \"split52_products_per_s\": %.4e", split);
    ms = time_ms([&] { k_wide<16><<<blocks, threads>>>((uint64_t*)out, 12345u, 678u, iters); });
    const double wide = total * 16 / (ms * 1e-3);
    printf(", \"imad_wide_per_s\": %.4e, \"imad_wide_lanes_per_clk_per_sm\": %.1f", wide, wide / sms / (khz * 1e3));
    // bits^2 of integer product per second: 52 x 52 per split product, 32 x 32 per IMAD.WIDE
    printf(", \"split52_bits2_per_s\": %.4e, \"imad_wide_bits2_per_s\": %.4e, \"ratio_split52_over_imad\": %.3f", split * 2704.0,
           wide * 1024.0, split * 2704.0 / (wide * 1024.0));
    ms = time_ms([&] { k_mixed<8, 8><<<blocks, threads>>>((double*)out, 1.5, 2.5, 99u, iters); });
    const double mixed_each = total * 8 / (ms * 1e-3);
    printf(", \"mixed_dfma_per_s\": %.4e, \"mixed_imad_wide_per_s\": %.4e", mixed_each, mixed_each);
    ms = time_ms([&] { k_mixed<16, 8><<<blocks, threads>>>((double*)out, 1.5, 2.5, 99u, iters); });
    printf(", \"mixed_2to1_dfma_per_s\": %.4e, \"mixed_2to1_imad_wide_per_s\": %.4e", total * 16 / (ms * 1e-3), total * 8 / (ms * 1e-3));
    printf("}\n");
    return 0;
}
