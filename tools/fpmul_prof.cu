// ncu target: the Montgomery multiply chain alone (development microbenchmark)
#include <cstdio>
#include <cuda_runtime.h>
#include "../mathlib_b200/csrc/curves.cuh"
using namespace b200;
template <class C>
__global__ void k_fpmul(Fp<C::N>* io, int iters) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    Fp<C::N> x = io[t], y = io[t];
    y.l[0] ^= 1;
    for (int i = 0; i < iters; i++) { FpOps<C>::mul(x, x, y); }
    io[t] = x;
}
int main() {
    void* buf; cudaMalloc(&buf, 148 * 4 * 128 * 48); cudaMemset(buf, 1, 148 * 4 * 128 * 48);
    for (int r = 0; r < 3; r++) k_fpmul<BLS381><<<148 * 4, 128>>>((Fp<12>*)buf, 400);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}
