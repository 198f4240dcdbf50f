// development check: the library's g1_mul2_kernel on one golden case, device vs the same headers on the host
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../mathlib_b200/csrc/kernels.cuh"
using namespace b200;
typedef BLS381 C;
typedef G1Ops<C> G;
typedef Codec<C> CD;
static std::vector<uint8_t> unhex(const char* s) {
    std::vector<uint8_t> o; size_t n = strlen(s);
    for (size_t i = 0; i + 1 < n; i += 2) { unsigned v; sscanf(s + i, "%2x", &v); o.push_back((uint8_t)v); }
    return o;
}
struct Prep { G::GlvTable t; uint32_t k1[5], k2[5]; };
__device__ __noinline__ void prep(Prep& p, const G::Aff& base, const uint32_t* k) {
    G::glv_split(p.k1, p.k2, k);
    G::glv_table(p.t, base);
}
template <int V>
__global__ void variant(const uint8_t* P, const uint8_t* es, const uint8_t* Q, const uint8_t* fs, uint8_t* out) {
    int err = 0; G::Aff a, b; uint32_t ke[8], kf[8];
    CD::g1_load(a.x, a.y, P, false, &err); CD::g1_load(b.x, b.y, Q, false, &err);
    CD::scalar_load(ke, es); CD::scalar_load(kf, fs);
    G::Pt acc;
    if (V == 1) {
        G::Pt acc2;
        G::scalar_mul(acc, a, ke); G::scalar_mul(acc2, b, kf); G::add(acc, acc2);
    } else if (V == 2) {
        Prep pp, pq;
        prep(pp, a, ke); prep(pq, b, kf);
        G::set_inf(acc);
        for (int i = 159; i >= 0; i--) {
            G::dbl(acc);
            const uint32_t bp = G::glv_bits(pp.k1, pp.k2, i), bq = G::glv_bits(pq.k1, pq.k2, i);
            if (bp) G::glv_step(acc, pp.t, bp);
            if (bq) G::glv_step(acc, pq.t, bq);
        }
    } else if (V == 3) {
        uint32_t e1[5], e2[5], f1[5], f2[5];
        G::glv_split(e1, e2, ke); G::glv_split(f1, f2, kf);
        G::GlvTable tp, tq;
        G::glv_table(tp, a); G::glv_table(tq, b);
        printf("dev f1 %08x %08x %08x %08x %08x f2 %08x %08x %08x %08x %08x tq.x2[0] %08x tq.x3[0] %08x\n", f1[0], f1[1], f1[2], f1[3], f1[4],
               f2[0], f2[1], f2[2], f2[3], f2[4], tq.x2.l[0], tq.x3.l[0]);
        G::scalar_mul2(acc, a, ke, b, kf);
    }
    G::to_affine(a, acc);
    CD::g1_store(out, a.x, a.y, false);
}
int main(int argc, char** argv) {
    auto p = unhex(argv[1]), e = unhex(argv[2]), q = unhex(argv[3]), f = unhex(argv[4]);
    int nthreads = argc > 5 ? atoi(argv[5]) : 1;
    // host
    int err = 0; G::Aff a, b; uint32_t ke[8], kf[8];
    CD::g1_load(a.x, a.y, p.data(), false, &err); CD::g1_load(b.x, b.y, q.data(), false, &err);
    CD::scalar_load(ke, e.data()); CD::scalar_load(kf, f.data());
    G::Pt acc; G::scalar_mul2(acc, a, ke, b, kf); G::to_affine(a, acc);
    uint8_t hout[96]; CD::g1_store(hout, a.x, a.y, false);
    // device, n copies
    size_t n = nthreads;
    uint8_t *dp, *de, *dq, *df, *dout; int* derr;
    cudaMalloc(&dp, 96 * n); cudaMalloc(&dq, 96 * n); cudaMalloc(&de, 32 * n); cudaMalloc(&df, 32 * n); cudaMalloc(&dout, 96 * n); cudaMalloc(&derr, 4);
    cudaMemset(derr, 0, 4);
    for (size_t i = 0; i < n; i++) {
        cudaMemcpy(dp + 96 * i, p.data(), 96, cudaMemcpyHostToDevice); cudaMemcpy(dq + 96 * i, q.data(), 96, cudaMemcpyHostToDevice);
        cudaMemcpy(de + 32 * i, e.data(), 32, cudaMemcpyHostToDevice); cudaMemcpy(df + 32 * i, f.data(), 32, cudaMemcpyHostToDevice);
    }
    g1_mul2_kernel<C><<<(unsigned)((n + 127) / 128), 128>>>(n, dp, de, dq, df, dout, 0, derr);
    std::vector<uint8_t> gout(96 * n);
    cudaMemcpy(gout.data(), dout, 96 * n, cudaMemcpyDeviceToHost);
    printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
    int bad = 0;
    for (size_t i = 0; i < n; i++) bad += memcmp(gout.data() + 96 * i, hout, 96) != 0;
    printf("mismatching threads: %d of %zu\nhost ", bad, n);
    for (int i = 0; i < 16; i++) printf("%02x", hout[i]);
    printf("\ndev  ");
    for (int i = 0; i < 16; i++) printf("%02x", gout[i]);
    printf("\n");
    {
        uint32_t f1[5], f2[5]; G::GlvTable tq; G::Aff bb; int er = 0;
        CD::g1_load(bb.x, bb.y, q.data(), false, &er);
        G::glv_split(f1, f2, kf); G::glv_table(tq, bb);
        printf("host f1 %08x %08x %08x %08x %08x f2 %08x %08x %08x %08x %08x tq.x2[0] %08x tq.x3[0] %08x\n", f1[0], f1[1], f1[2], f1[3], f1[4],
               f2[0], f2[1], f2[2], f2[3], f2[4], tq.x2.l[0], tq.x3.l[0]);
    }
    for (int v = 1; v <= 3; v++) {
        if (v == 1) variant<1><<<1, 1>>>(dp, de, dq, df, dout);
        if (v == 2) variant<2><<<1, 1>>>(dp, de, dq, df, dout);
        if (v == 3) variant<3><<<1, 1>>>(dp, de, dq, df, dout);
        cudaMemcpy(gout.data(), dout, 96, cudaMemcpyDeviceToHost);
        printf("variant %d: %s  %s\n", v, memcmp(gout.data(), hout, 96) ? "MISMATCH" : "ok", cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
