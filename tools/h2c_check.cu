// Development aid: run the stages of hash_to_g1.cuh on the device and on the host (same headers) and report the first
// stage whose outputs differ.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/h2c_check.bin tools/h2c_check.cu
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include "../mathlib_b200/csrc/hash_to_g1.cuh"
using namespace b200;
typedef HashToG1 H;
struct Stages { uint8_t ub[128]; uint32_t u0[12], u1[12], x0[12], y0[12], x1[12], y1[12], xs[12], ys[12], qx[12], qy[12]; int ok_add, ok_iso; uint32_t cx[12], cy[12]; uint8_t out[96]; };
template <class HF>
B200_HD void stages(Stages* s, const uint8_t* msg, size_t mlen, const uint8_t* dst, size_t dlen, bool be) {
    H::expand128<HF>(s->ub, msg, mlen, dst, dlen);
    H::E u0, u1, x0, y0, x1, y1, xs, ys, qx, qy;
    H::field_from_64(u0, s->ub); H::field_from_64(u1, s->ub + 64);
    H::swu(x0, y0, u0, be); H::swu(x1, y1, u1, be);
    s->ok_add = H::add_iso(xs, ys, x0, y0, x1, y1);
    s->ok_iso = H::iso_map(qx, qy, xs, ys);
    for (int i = 0; i < 12; i++) { s->u0[i] = u0.l[i]; s->u1[i] = u1.l[i]; s->x0[i] = x0.l[i]; s->y0[i] = y0.l[i]; s->x1[i] = x1.l[i]; s->y1[i] = y1.l[i];
        s->xs[i] = xs.l[i]; s->ys[i] = ys.l[i]; s->qx[i] = qx.l[i]; s->qy[i] = qy.l[i]; }
    { G1Ops<BLS381>::Aff q, r; q.x = qx; q.y = qy; H::clear_cofactor(r, q); for (int i = 0; i < 12; i++) { s->cx[i] = r.x.l[i]; s->cy[i] = r.y.l[i]; } }
    H::item(be ? 1 : 0, msg, mlen, dst, dlen, s->out, false);
}
__global__ void k(Stages* s, const uint8_t* msg, size_t mlen, const uint8_t* dst, size_t dlen, int be) {
    if (be) stages<Blake2b512>(s, msg, mlen, dst, dlen, true); else stages<Sha256>(s, msg, mlen, dst, dlen, false);
}
int main() {
    const char* msgs[3] = {"Chase!", "", "abc"};
    int bad = 0;
    for (int be = 0; be < 2; be++) for (int m = 0; m < 3; m++) {
        size_t mlen = strlen(msgs[m]);
        Stages h, d, *dd; uint8_t* dm;
        memset(&h, 0, sizeof h);
        if (be) stages<Blake2b512>(&h, (const uint8_t*)msgs[m], mlen, (const uint8_t*)"EF", 2, true);
        else stages<Sha256>(&h, (const uint8_t*)msgs[m], mlen, (const uint8_t*)"EF", 2, false);
        cudaMalloc(&dd, sizeof(Stages)); cudaMalloc(&dm, 64); cudaMemset(dd, 0, sizeof(Stages));
        cudaMemcpy(dm, msgs[m], mlen, cudaMemcpyHostToDevice); cudaMemcpy(dm + 32, "EF", 2, cudaMemcpyHostToDevice);
        k<<<1, 1>>>(dd, dm, mlen, dm + 32, 2, be);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&d, dd, sizeof(Stages), cudaMemcpyDeviceToHost);
        printf("be=%d msg=%d cuda=%s:", be, m, cudaGetErrorString(e));
#define CMP(f) if (memcmp(&h.f, &d.f, sizeof h.f)) { printf(" %s DIFFERS", #f); bad++; } else printf(" %s ok", #f);
        CMP(ub) CMP(u0) CMP(u1) CMP(x0) CMP(y0) CMP(x1) CMP(y1) CMP(ok_add) CMP(xs) CMP(ys) CMP(ok_iso) CMP(qx) CMP(qy) CMP(cx) CMP(cy) CMP(out)
        printf("\n");
    }
    return bad ? 1 : 0;
}
