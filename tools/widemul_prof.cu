// development microbenchmark: does the VM's wide (unreduced) N x N product saturate the IMAD.WIDE pipe, and with how many
// warps per SM sub-partition?  Prints multiply-accumulates per second for 1..4 warps per sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
#include "../mathlib_b200/csrc/vm.cuh"
using namespace b200;
typedef Vm<BLS381> M;
__global__ void k(uint32_t* io, int iters) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t a[12], b[12], v[24];
    for (int i = 0; i < 12; i++) { a[i] = io[t * 24 + i]; b[i] = io[t * 24 + 12 + i]; }
    for (int it = 0; it < iters; it++) {
        M::wide_mul(v, a, b);
        for (int i = 0; i < 12; i++) { a[i] ^= v[i]; b[i] += v[12 + i]; }
    }
    for (int i = 0; i < 12; i++) io[t * 24 + i] = a[i] ^ b[i];
}
int main() {
    const int sms = 148, iters = 2000;
    uint32_t* buf; cudaMalloc(&buf, (size_t)sms * 1024 * 24 * 4); cudaMemset(buf, 7, (size_t)sms * 1024 * 24 * 4);
    for (int warps_per_smsp = 1; warps_per_smsp <= 4; warps_per_smsp++) {
        int threads = warps_per_smsp * 128;
        k<<<sms, threads>>>(buf, 10); cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); k<<<sms, threads>>>(buf, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double macs = (double)sms * threads * iters * 144.0;
        printf("{\"warps_per_smsp\": %d, \"mac_per_s\": %.4g, \"frac_of_8.98T\": %.3f}\n", warps_per_smsp, macs / (ms * 1e-3), macs / (ms * 1e-3) / 8.98e12);
    }
    return 0;
}
