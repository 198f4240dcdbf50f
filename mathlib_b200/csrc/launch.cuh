// Host-side launch wrappers, instantiated once per curve (kernels_<curve>.cu) and reached through a
// small table of function pointers from the C ABI (abi.cu).  No CPU compute path exists here: every
// entry launches sm_100a kernels on the given stream.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdlib>
#include <map>
#include <mutex>
#include <type_traits>
#include "msm.cuh"
#include "pairing_vm.cuh"
#include "g2.cuh"
#include "points.cuh"
#include "g1_p3.cuh"
#include "hash_to_g1.cuh"

namespace b200 {

extern std::atomic<uint64_t> g_launch_count;

struct MsmBuffers {            // device pointers carved out of one workspace slab by the ABI layer
    void* points;              // n * sizeof(G1Affine)   (unused when points are already resident)
    uint32_t* digits;          // W * n
    uint32_t* sorted;          // W * n
    uint32_t* counts;          // W * B
    uint32_t* offsets;         // W * B
    uint32_t* cursor;          // W * B
    uint32_t* perm;            // W * B   bucket ids by decreasing size
    uint32_t* size_hist;       // 2 * B200_MSM_SIZE_BINS (histogram, start offsets)
    void* buckets;             // W * B * sizeof(G1XYZZ)
    void* chunks;              // W * nchunks * sizeof(G1XYZZ)
    void* windows;             // W * sizeof(G1XYZZ)
    uint32_t* heavy_n;         // number of heavy items (device counter)
    void* heavy_items;         // max_heavy * sizeof(MsmHeavyItem)
    void* heavy_partial;       // max_heavy * sizeof(G1XYZZ)
    uint32_t max_heavy;        // W * n / B200_MSM_SEG
    cudaEvent_t points_ready = nullptr;   // if set: the bucket kernels wait for it (points prepared on another stream)
};

struct CurveVTable {
    int fp_bytes;
    int limbs;
    int scalar_bits;
    size_t aff_size;           // sizeof(G1Affine<N>)
    size_t xyzz_size;          // sizeof(G1XYZZ<N>)
    int glv;                   // > 0: bit length of the halves of an exact GLV split (BLS12 family): one-shot MSMs split the scalars
    cudaError_t (*pairing)(int np, size_t n, const uint8_t* g1a, const uint8_t* g2a, const uint8_t* g1b,
                           const uint8_t* g2b, uint8_t* out, uint32_t flags, int* err, cudaStream_t s);
    cudaError_t (*fexp)(size_t n, const uint8_t* in, uint8_t* out, uint32_t flags, int* err, cudaStream_t s);
    cudaError_t (*g1_mul)(size_t n, const uint8_t* pts, const uint8_t* k, uint8_t* out, uint32_t flags, int* err,
                          cudaStream_t s);
    cudaError_t (*g1_mul2)(size_t n, const uint8_t* P, const uint8_t* e, const uint8_t* Q, const uint8_t* f,
                           uint8_t* out, uint32_t flags, int* err, cudaStream_t s);
    cudaError_t (*g1_sum)(size_t n, const uint8_t* pts, uint8_t* out, uint32_t flags, int* err, cudaStream_t s);
    // G2.Mul / G2.Add batches, Gt.Mul / Gt.Inverse / Gt.Exp batches (op = GT_OP_*)
    cudaError_t (*g2_mul)(size_t n, const uint8_t* pts, const uint8_t* k, uint8_t* out, uint32_t flags, int* err,
                          cudaStream_t s);
    cudaError_t (*g2_sum)(size_t n, const uint8_t* pts, uint8_t* out, uint32_t flags, int* err, cudaStream_t s);
    cudaError_t (*gt_op)(int op, size_t n, const uint8_t* a, const uint8_t* b, uint8_t* out, uint32_t flags, int* err,
                         cudaStream_t s);
    // fixed-Q line tables -- rows of lines_row_words words per G2 point; pairings against table rows
    size_t (*lines_row_words)();
    cudaError_t (*lines_build)(size_t n_q, const uint8_t* g2, uint32_t* lines, uint8_t* qinf, uint32_t flags, int* err,
                               cudaStream_t s);
    cudaError_t (*pairing_fixed)(int np, size_t n, const uint8_t* g1a, const uint32_t* qa_idx, const uint8_t* g1b,
                                 const uint32_t* qb_idx, const uint32_t* lines, const uint8_t* qinf, uint32_t n_q,
                                 uint8_t* out, uint32_t flags, int* err, cudaStream_t s);
    // point decompression (op 0) / compression (1) / validation (2) batches, g2 = 0 / 1
    cudaError_t (*point_codec)(int g2, int op, size_t n, const uint8_t* in, uint8_t* out, uint32_t flags, int* err,
                               cudaStream_t s);
    // batch affine normalisation of Jacobian G1 points (Montgomery's trick)
    cudaError_t (*g1_normalize)(size_t n, const uint32_t* jac, uint8_t* out, uint32_t flags, cudaStream_t s);
    // hash-to-G1 batch (BLS12-381 only; nullptr on the other curves)
    cudaError_t (*hash_to_g1)(int bbs, size_t n, const uint8_t* msgs, const uint64_t* offsets, const uint8_t* dst, size_t dlen,
                              uint8_t* out, uint32_t flags, cudaStream_t s);
    // points -> Montgomery affine array (for MSM / resident bases)
    // glv != 0: out holds 2n points, P_i and phi(P_i) (MsmPlan.glv)
    cudaError_t (*msm_points)(size_t n, const uint8_t* pts, void* out, uint32_t flags, int* err, cudaStream_t s, int glv);
    // resident window tables: tab[w*stride + i] = 2^(c*w) * tab[i]
    cudaError_t (*msm_tables)(size_t n, const MsmPlan& pl, size_t stride, void* tab, cudaStream_t s);
    // MSM over prepared points
    cudaError_t (*msm)(size_t n, const void* prepared_pts, const uint8_t* scalars, uint8_t* out, uint32_t flags,
                       const MsmPlan& pl, const MsmBuffers& b, cudaStream_t s);
    // G2 MSM: BYTES -> Montgomery affine G2 points, then the same pipeline over Fp2
    cudaError_t (*msm_points_g2)(size_t n, const uint8_t* pts, void* out, uint32_t flags, int* err, cudaStream_t s);
    cudaError_t (*msm_g2)(size_t n, const void* prepared_pts, const uint8_t* scalars, uint8_t* out, uint32_t flags,
                          const MsmPlan& pl, const MsmBuffers& b, cudaStream_t s);
};

const CurveVTable* vtable_bn254();
const CurveVTable* vtable_bls381();
const CurveVTable* vtable_bls377();

#if defined(B200_INSTANTIATE)
static inline unsigned blocks_for(size_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }
#define B200_COUNT_LAUNCH() g_launch_count.fetch_add(1, std::memory_order_relaxed)

template <class C>
struct Launch {
    // development knob (-DB200_DEV_KNOBS builds only): unused dynamic shared memory caps the resident blocks of the G1
    // kernels (measured: no gain)
    static size_t g1_smem_pad() {
#if defined(B200_DEV_KNOBS)
        static long v = getenv("B200_G1_SMEM_PAD") ? atol(getenv("B200_G1_SMEM_PAD")) : 0;
        return (size_t)v;
#else
        return 0;
#endif
    }
    static cudaError_t pairing(int np, size_t n, const uint8_t* g1a, const uint8_t* g2a, const uint8_t* g1b,
                               const uint8_t* g2b, uint8_t* out, uint32_t flags, int* err, cudaStream_t s) {
        if (n == 0) return cudaSuccess;
#if defined(B200_PAIR_VARIANTS)
        // development switch: occupancy / register-budget variants of the same kernel
        static int variant = getenv("B200_PAIR_VARIANT") ? atoi(getenv("B200_PAIR_VARIANT")) : 0;
#define B200_LAUNCH_VARIANT(T, MB)                                                                              \
        do {                                                                                                    \
            unsigned nbv = blocks_for(n, T);                                                                    \
            if (np == 1) pairing_kernel<C, 1, T, MB><<<nbv, T, 0, s>>>(n, g1a, g2a, g1a, g2a, out, flags, err); \
            else pairing_kernel<C, 2, T, MB><<<nbv, T, 0, s>>>(n, g1a, g2a, g1b, g2b, out, flags, err);         \
        } while (0)
        if (variant == 1) { B200_LAUNCH_VARIANT(128, 3); B200_COUNT_LAUNCH(); return cudaGetLastError(); }
        if (variant == 2) { B200_LAUNCH_VARIANT(128, 4); B200_COUNT_LAUNCH(); return cudaGetLastError(); }
        if (variant == 3) { B200_LAUNCH_VARIANT(256, 3); B200_COUNT_LAUNCH(); return cudaGetLastError(); }
        if (variant == 4) { B200_LAUNCH_VARIANT(256, 4); B200_COUNT_LAUNCH(); return cudaGetLastError(); }
#endif
#if defined(B200_PAIR_LEGACY_BUILD)
        // thread-per-pairing kernel (pairing.cuh): an independent cross-check of the VM kernel, only in builds made with
        // -DB200_PAIR_LEGACY_BUILD (it is 600 KB of code and a minute of compile time; not in the product library)
        // (65,536 BN254 Pairing+FExp: 63.5 ms here vs 58.7 ms on the VM; BLS12-381 Pairing2+FExp: 128 ms vs 100 ms)
        static int legacy = getenv("B200_PAIR_LEGACY") ? atoi(getenv("B200_PAIR_LEGACY")) : 0;
        if (legacy) {
            unsigned nb = blocks_for(n, B200_PAIR_THREADS);
            if (np == 1) pairing_kernel<C, 1><<<nb, B200_PAIR_THREADS, 0, s>>>(n, g1a, g2a, g1a, g2a, out, flags, err);
            else pairing_kernel<C, 2><<<nb, B200_PAIR_THREADS, 0, s>>>(n, g1a, g2a, g1b, g2b, out, flags, err);
            B200_COUNT_LAUNCH();
            return cudaGetLastError();
        }
#endif
        return vm_pairing(np, n, g1a, g2a, g1b, g2b, out, flags, err, s);
    }
    // at most one check per resident warp (148 SMs x 2 blocks x 4 warps): the split kernel's single wave.  Beyond that a
    // second wave of the split kernel costs as much as one wave of the five-checks-per-warp kernel.
    static size_t split_max() {
#if defined(B200_DEV_KNOBS)
        static long v = getenv("B200_VM_SPLIT_MAX") ? atol(getenv("B200_VM_SPLIT_MAX")) : 148 * 2 * B200_VM_SPLIT_WARPS;
        return (size_t)v;
#else
        return 148 * 2 * B200_VM_SPLIT_WARPS;
#endif
    }
    // batches that cannot fill every SM with full-size blocks use the 4-warp variant (20 products per block)
    static bool small_batch(size_t n) { return n < (size_t)148 * vm_warps<C>() * B200_VM_GROUPS_PER_WARP; }
    template <int W>
    static void vm_launch(int np, size_t n, const uint8_t* g1a, const uint8_t* g2a, const uint8_t* g1b, const uint8_t* g2b,
                          uint8_t* out, uint32_t flags, int* err, const uint32_t* d_words, const VmDirEntry* d_dir,
                          cudaStream_t s) {
        const unsigned gpb = W * B200_VM_GROUPS_PER_WARP;
        const unsigned nb = (unsigned)((n + gpb - 1) / gpb);
        const size_t smem = vm_smem_bytes<C, W, true>();
        if (np == 1)
            vm_pairing_kernel<C, 1, W><<<nb, W * 32, smem, s>>>(n, g1a, g2a, g1a, g2a, out, flags, err, d_words, d_dir);
        else
            vm_pairing_kernel<C, 2, W><<<nb, W * 32, smem, s>>>(n, g1a, g2a, g1b, g2b, out, flags, err, d_words, d_dir);
    }
    // per device: upload the microcode once and opt in to the large dynamic shared-memory carve-out
    static cudaError_t vm_setup(const uint32_t** words, const VmDirEntry** dir) {
        static thread_local int cfg_dev = -1;
        static thread_local const uint32_t* d_words = nullptr;
        static thread_local const VmDirEntry* d_dir = nullptr;
        cudaError_t e;
        int dev = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
        if (cfg_dev != dev) {
            static std::mutex mu;
            static std::map<int, std::pair<const uint32_t*, const VmDirEntry*>> tables;
            std::lock_guard<std::mutex> lk(mu);
            auto it = tables.find(dev);
            if (it == tables.end()) {
                uint32_t* w = nullptr;
                VmDirEntry* d = nullptr;
                if ((e = cudaMalloc(&w, sizeof(uint32_t) * VmTables<C>::NWORDS)) != cudaSuccess) return e;
                if ((e = cudaMalloc(&d, sizeof(VmDirEntry) * VP_COUNT)) != cudaSuccess) return e;
                if ((e = cudaMemcpy(w, VmTables<C>::host_words(), sizeof(uint32_t) * VmTables<C>::NWORDS,
                                    cudaMemcpyHostToDevice)) != cudaSuccess) return e;
                if ((e = cudaMemcpy(d, VmTables<C>::host_dir(), sizeof(VmDirEntry) * VP_COUNT, cudaMemcpyHostToDevice)) !=
                    cudaSuccess) return e;
                constexpr int WB = vm_warps<C>(), WS = B200_VM_WARPS_SMALL;
                constexpr int WX = vm_warps_x<C>();
                const int sb = (int)vm_smem_bytes<C, WX>(), ss = (int)vm_smem_bytes<C, WS>();
                const int cb = (int)vm_smem_bytes<C, WB, true>(), cs = (int)vm_smem_bytes<C, WS, true>();
                const cudaFuncAttribute at = cudaFuncAttributeMaxDynamicSharedMemorySize;
                if ((e = cudaFuncSetAttribute(vm_pairing_kernel<C, 1, WB>, at, cb)) != cudaSuccess) return e;
                if ((e = cudaFuncSetAttribute(vm_pairing_kernel<C, 2, WB>, at, cb)) != cudaSuccess) return e;
                if ((e = cudaFuncSetAttribute(vm_fexp_kernel<C, WB>, at, cb)) != cudaSuccess) return e;
                if ((e = cudaFuncSetAttribute(vm_pairing_kernel<C, 1, WS>, at, cs)) != cudaSuccess) return e;
                if ((e = cudaFuncSetAttribute(vm_pairing_kernel<C, 2, WS>, at, cs)) != cudaSuccess) return e;
                if ((e = cudaFuncSetAttribute(vm_fexp_kernel<C, WS>, at, cs)) != cudaSuccess) return e;
                if ((e = cudaFuncSetAttribute(vm_pairing_split_kernel<C, 1>, at, (int)vm_split_smem_bytes<C>())) != cudaSuccess) return e;
                if ((e = cudaFuncSetAttribute(vm_pairing_split_kernel<C, 2>, at, (int)vm_split_smem_bytes<C>())) != cudaSuccess) return e;
                if ((e = cudaFuncSetAttribute(vm_lines_kernel<C, WS>, at, ss)) != cudaSuccess) return e;
                if ((e = cudaFuncSetAttribute(vm_pairing_fixed_kernel<C, 1, WX>, at, sb)) != cudaSuccess) return e;
                if ((e = cudaFuncSetAttribute(vm_pairing_fixed_kernel<C, 2, WX>, at, sb)) != cudaSuccess) return e;
                if ((e = cudaFuncSetAttribute(vm_pairing_fixed_kernel<C, 1, WS>, at, ss)) != cudaSuccess) return e;
                if ((e = cudaFuncSetAttribute(vm_pairing_fixed_kernel<C, 2, WS>, at, ss)) != cudaSuccess) return e;
                if ((e = cudaFuncSetAttribute(vm_gt_kernel<C, WX>, at, sb)) != cudaSuccess) return e;
                if ((e = cudaFuncSetAttribute(vm_gt_kernel<C, WS>, at, ss)) != cudaSuccess) return e;
                it = tables.emplace(dev, std::make_pair((const uint32_t*)w, (const VmDirEntry*)d)).first;
            }
            d_words = it->second.first;
            d_dir = it->second.second;
            cfg_dev = dev;
        }
        *words = d_words;
        *dir = d_dir;
        return cudaSuccess;
    }
    // warp-cooperative VM kernel (pairing_vm.cuh): 6 lanes per pairing product, 40 products per 256-thread block
    static cudaError_t vm_pairing(int np, size_t n, const uint8_t* g1a, const uint8_t* g2a, const uint8_t* g1b,
                                  const uint8_t* g2b, uint8_t* out, uint32_t flags, int* err, cudaStream_t s) {
        const uint32_t* d_words = nullptr;
        const VmDirEntry* d_dir = nullptr;
        cudaError_t e = vm_setup(&d_words, &d_dir);
        if (e != cudaSuccess) return e;
        if (n <= split_max()) {
            // fewer checks than resident warps (BASELINE configs[0]): one check per warp, three lanes per role
            const unsigned nb = (unsigned)((n + B200_VM_SPLIT_WARPS - 1) / B200_VM_SPLIT_WARPS);
            const size_t smem = vm_split_smem_bytes<C>();
            if (np == 1)
                vm_pairing_split_kernel<C, 1><<<nb, B200_VM_SPLIT_WARPS * 32, smem, s>>>(n, g1a, g2a, g1a, g2a, out, flags, err, d_words, d_dir);
            else
                vm_pairing_split_kernel<C, 2><<<nb, B200_VM_SPLIT_WARPS * 32, smem, s>>>(n, g1a, g2a, g1b, g2b, out, flags, err, d_words, d_dir);
        } else if (small_batch(n)) vm_launch<B200_VM_WARPS_SMALL>(np, n, g1a, g2a, g1b, g2b, out, flags, err, d_words, d_dir, s);
        else vm_launch<vm_warps<C>()>(np, n, g1a, g2a, g1b, g2b, out, flags, err, d_words, d_dir, s);
        B200_COUNT_LAUNCH();
        return cudaGetLastError();
    }
    static cudaError_t fexp(size_t n, const uint8_t* in, uint8_t* out, uint32_t flags, int* err, cudaStream_t s) {
        if (n == 0) return cudaSuccess;
#if defined(B200_PAIR_LEGACY_BUILD)
        static int legacy = getenv("B200_PAIR_LEGACY") ? atoi(getenv("B200_PAIR_LEGACY")) : 0;
        if (legacy) {
            fexp_kernel<C><<<blocks_for(n, B200_PAIR_THREADS), B200_PAIR_THREADS, 0, s>>>(n, in, out, flags, err);
            B200_COUNT_LAUNCH();
            return cudaGetLastError();
        }
#endif
        const uint32_t* d_words = nullptr;
        const VmDirEntry* d_dir = nullptr;
        cudaError_t e = vm_setup(&d_words, &d_dir);
        if (e != cudaSuccess) return e;
        if (small_batch(n)) {
            constexpr int W = B200_VM_WARPS_SMALL;
            const unsigned gpb = W * B200_VM_GROUPS_PER_WARP;
            vm_fexp_kernel<C, W><<<(unsigned)((n + gpb - 1) / gpb), W * 32, vm_smem_bytes<C, W, true>(), s>>>(n, in, out, flags,
                                                                                                          err, d_words, d_dir);
        } else {
            constexpr int W = vm_warps<C>();
            const unsigned gpb = W * B200_VM_GROUPS_PER_WARP;
            vm_fexp_kernel<C, W><<<(unsigned)((n + gpb - 1) / gpb), W * 32, vm_smem_bytes<C, W, true>(), s>>>(n, in, out, flags,
                                                                                                          err, d_words, d_dir);
        }
        B200_COUNT_LAUNCH();
        return cudaGetLastError();
    }
    static size_t lines_row_words() { return (size_t)VmDriver<C>::nlines() * 3 * 2 * C::N; }
    static cudaError_t lines_build(size_t n_q, const uint8_t* g2, uint32_t* lines, uint8_t* qinf, uint32_t flags, int* err,
                                   cudaStream_t s) {
        if (n_q == 0) return cudaSuccess;
        const uint32_t* d_words = nullptr;
        const VmDirEntry* d_dir = nullptr;
        cudaError_t e = vm_setup(&d_words, &d_dir);
        if (e != cudaSuccess) return e;
        constexpr int W = B200_VM_WARPS_SMALL;
        const unsigned gpb = W * B200_VM_GROUPS_PER_WARP;
        vm_lines_kernel<C, W><<<(unsigned)((n_q + gpb - 1) / gpb), W * 32, vm_smem_bytes<C, W>(), s>>>(n_q, g2, lines, qinf, flags,
                                                                                                   err, d_words, d_dir);
        B200_COUNT_LAUNCH();
        return cudaGetLastError();
    }
    template <int W>
    static void fixed_launch(int np, size_t n, const uint8_t* g1a, const uint32_t* qa, const uint8_t* g1b, const uint32_t* qb,
                             const uint32_t* lines, const uint8_t* qinf, uint32_t n_q, uint8_t* out, uint32_t flags, int* err,
                             const uint32_t* d_words, const VmDirEntry* d_dir, cudaStream_t s) {
        const unsigned gpb = W * B200_VM_GROUPS_PER_WARP;
        const unsigned nb = (unsigned)((n + gpb - 1) / gpb);
        const size_t smem = vm_smem_bytes<C, W>();
        if (np == 1)
            vm_pairing_fixed_kernel<C, 1, W><<<nb, W * 32, smem, s>>>(n, g1a, qa, g1a, qa, lines, qinf, n_q, out, flags, err, d_words, d_dir);
        else
            vm_pairing_fixed_kernel<C, 2, W><<<nb, W * 32, smem, s>>>(n, g1a, qa, g1b, qb, lines, qinf, n_q, out, flags, err, d_words, d_dir);
    }
    static cudaError_t pairing_fixed(int np, size_t n, const uint8_t* g1a, const uint32_t* qa, const uint8_t* g1b,
                                     const uint32_t* qb, const uint32_t* lines, const uint8_t* qinf, uint32_t n_q,
                                     uint8_t* out, uint32_t flags, int* err, cudaStream_t s) {
        if (n == 0) return cudaSuccess;
        const uint32_t* d_words = nullptr;
        const VmDirEntry* d_dir = nullptr;
        cudaError_t e = vm_setup(&d_words, &d_dir);
        if (e != cudaSuccess) return e;
        if (small_batch(n)) fixed_launch<B200_VM_WARPS_SMALL>(np, n, g1a, qa, g1b, qb, lines, qinf, n_q, out, flags, err, d_words, d_dir, s);
        else fixed_launch<vm_warps_x<C>()>(np, n, g1a, qa, g1b, qb, lines, qinf, n_q, out, flags, err, d_words, d_dir, s);
        B200_COUNT_LAUNCH();
        return cudaGetLastError();
    }
    static cudaError_t gt_op(int op, size_t n, const uint8_t* a, const uint8_t* b, uint8_t* out, uint32_t flags, int* err,
                             cudaStream_t s) {
        if (n == 0) return cudaSuccess;
        const uint32_t* d_words = nullptr;
        const VmDirEntry* d_dir = nullptr;
        cudaError_t e = vm_setup(&d_words, &d_dir);
        if (e != cudaSuccess) return e;
        if (small_batch(n)) {
            constexpr int W = B200_VM_WARPS_SMALL;
            const unsigned gpb = W * B200_VM_GROUPS_PER_WARP;
            vm_gt_kernel<C, W><<<(unsigned)((n + gpb - 1) / gpb), W * 32, vm_smem_bytes<C, W>(), s>>>(op, n, a, b, out, flags,
                                                                                                  err, d_words, d_dir);
        } else {
            constexpr int W = vm_warps_x<C>();
            const unsigned gpb = W * B200_VM_GROUPS_PER_WARP;
            vm_gt_kernel<C, W><<<(unsigned)((n + gpb - 1) / gpb), W * 32, vm_smem_bytes<C, W>(), s>>>(op, n, a, b, out, flags,
                                                                                                  err, d_words, d_dir);
        }
        B200_COUNT_LAUNCH();
        return cudaGetLastError();
    }
    static cudaError_t point_codec(int g2, int op, size_t n, const uint8_t* in, uint8_t* out, uint32_t flags, int* err,
                                   cudaStream_t s) {
        if (n == 0) return cudaSuccess;
        if (g2) point_codec_kernel<C, 1><<<blocks_for(n, B200_PT_THREADS), B200_PT_THREADS, 0, s>>>(op, n, in, out, flags, err);
        else point_codec_kernel<C, 0><<<blocks_for(n, B200_PT_THREADS), B200_PT_THREADS, 0, s>>>(op, n, in, out, flags, err);
        B200_COUNT_LAUNCH();
        return cudaGetLastError();
    }
    static cudaError_t g1_normalize(size_t n, const uint32_t* jac, uint8_t* out, uint32_t flags, cudaStream_t s) {
        if (n == 0) return cudaSuccess;
        g1_normalize_kernel<C><<<blocks_for((n + B200_NORM_BATCH - 1) / B200_NORM_BATCH, 64), 64, 0, s>>>(n, jac, out, flags);
        B200_COUNT_LAUNCH();
        return cudaGetLastError();
    }
    static cudaError_t hash_to_g1(int bbs, size_t n, const uint8_t* msgs, const uint64_t* offsets, const uint8_t* dst,
                                  size_t dlen, uint8_t* out, uint32_t flags, cudaStream_t s) {
        if (n == 0) return cudaSuccess;
        hash_to_g1_kernel<<<blocks_for(n, 64), 64, 0, s>>>(bbs, n, msgs, offsets, dst, dlen, out, flags);
        B200_COUNT_LAUNCH();
        return cudaGetLastError();
    }
    static cudaError_t g2_mul(size_t n, const uint8_t* pts, const uint8_t* k, uint8_t* out, uint32_t flags, int* err,
                              cudaStream_t s) {
        if (n == 0) return cudaSuccess;
        g2_mul_kernel<C><<<blocks_for(n, B200_G2_THREADS), B200_G2_THREADS, 0, s>>>(n, pts, k, out, flags, err);
        B200_COUNT_LAUNCH();
        return cudaGetLastError();
    }
    static cudaError_t g2_sum(size_t n, const uint8_t* pts, uint8_t* out, uint32_t flags, int* err, cudaStream_t s) {
        g2_sum_kernel<C><<<1, 32, 0, s>>>(n, pts, out, flags, err);
        B200_COUNT_LAUNCH();
        return cudaGetLastError();
    }
    // batches up to this size run three lanes per operation (g1_p3.cuh).  Measured: 1,000 Mul 2.5 ms against 3.4 ms, but
    // 12,500 Mul 5.6 ms against 4.3 ms -- the exchange and the repeated additions cost more than the shorter chain saves
    // once the one-lane kernel has enough warps, so only really small calls take this path.
    static size_t p3_max() {
#if defined(B200_DEV_KNOBS)
        static long v = getenv("B200_G1_P3_MAX") ? atol(getenv("B200_G1_P3_MAX")) : 2048;
        return (size_t)v;
#else
        return 2048;
#endif
    }
    static cudaError_t g1_mul(size_t n, const uint8_t* pts, const uint8_t* k, uint8_t* out, uint32_t flags, int* err,
                              cudaStream_t s) {
        if (n == 0) return cudaSuccess;
        if (n <= p3_max()) {
            g1_mul_p3_kernel<C, 1><<<blocks_for(n, B200_P3_OPS_PER_BLOCK), B200_P3_THREADS, 0, s>>>(n, pts, k, pts, k, out, flags, err);
            B200_COUNT_LAUNCH();
            return cudaGetLastError();
        }
        g1_mul_kernel<C><<<blocks_for(n, B200_G1_THREADS), B200_G1_THREADS, g1_smem_pad(), s>>>(n, pts, k, out, flags, err);
        B200_COUNT_LAUNCH();
        return cudaGetLastError();
    }
    static cudaError_t g1_mul2(size_t n, const uint8_t* P, const uint8_t* e, const uint8_t* Q, const uint8_t* f,
                               uint8_t* out, uint32_t flags, int* err, cudaStream_t s) {
        if (n == 0) return cudaSuccess;
        if (n <= p3_max()) {
            g1_mul_p3_kernel<C, 2><<<blocks_for(n, B200_P3_OPS_PER_BLOCK), B200_P3_THREADS, 0, s>>>(n, P, e, Q, f, out, flags, err);
            B200_COUNT_LAUNCH();
            return cudaGetLastError();
        }
        g1_mul2_kernel<C><<<blocks_for(n, B200_G1_THREADS), B200_G1_THREADS, g1_smem_pad(), s>>>(n, P, e, Q, f, out, flags, err);
        B200_COUNT_LAUNCH();
        return cudaGetLastError();
    }
    static cudaError_t g1_sum(size_t n, const uint8_t* pts, uint8_t* out, uint32_t flags, int* err, cudaStream_t s) {
        g1_sum_kernel<C><<<1, 32, 0, s>>>(n, pts, out, flags, err);
        B200_COUNT_LAUNCH();
        return cudaGetLastError();
    }
    static cudaError_t msm_points(size_t n, const uint8_t* pts, void* out, uint32_t flags, int* err, cudaStream_t s, int glv) {
        if (n == 0) return cudaSuccess;
        msm_points_kernel<C><<<blocks_for(n, 128), 128, 0, s>>>(n, pts, (G1Affine<C::N>*)out, flags, err, glv);
        B200_COUNT_LAUNCH();
        return cudaGetLastError();
    }
    static cudaError_t msm_tables(size_t n, const MsmPlan& pl, size_t stride, void* tab, cudaStream_t s) {
        if (n == 0 || pl.W < 2) return cudaSuccess;
        msm_tables_kernel<C><<<blocks_for(n, 128), 128, 0, s>>>(n, pl, stride, (G1Affine<C::N>*)tab);
        B200_COUNT_LAUNCH();
        return cudaGetLastError();
    }
    // the Pippenger pipeline for G = G1Ops<C> (driver.Curve.MultiScalarMul) or G2Ops<C> (G2 MSM)
    template <class G>
    static cudaError_t msm_impl(size_t n, const void* pts, const uint8_t* scalars, uint8_t* out, uint32_t flags,
                                const MsmPlan& pl, const MsmBuffers& b, cudaStream_t s) {
        typedef typename G::Pt Pt;
        typedef typename G::Aff Aff;
        constexpr bool IS_G1 = std::is_same<G, G1Ops<C>>::value;
        size_t nb = (size_t)pl.W * pl.B;
        cudaError_t e;
        if ((e = cudaMemsetAsync(b.counts, 0, nb * sizeof(uint32_t), s)) != cudaSuccess) return e;
        if (n) {
            msm_digits_kernel<C><<<blocks_for(n, 256), 256, 0, s>>>(n, scalars, pl, b.digits, b.counts);
            B200_COUNT_LAUNCH();
        }
        if (pl.glv) n *= 2;                          // from here on: 2n points (P_i, phi(P_i)) with 130-bit scalars
        int st = pl.B >= 1024 ? 1024 : (pl.B >= 32 ? pl.B : 32);
        msm_scan_kernel<<<pl.W, st, st * sizeof(uint32_t), s>>>(pl, b.counts, b.offsets);
        B200_COUNT_LAUNCH();
        if ((e = cudaMemcpyAsync(b.cursor, b.offsets, nb * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s)) != cudaSuccess)
            return e;
        if (n) {
            msm_scatter_kernel<<<blocks_for(n, 256), 256, 0, s>>>(n, pl, b.digits, b.cursor, b.sorted);
            B200_COUNT_LAUNCH();
        }
        if (b.max_heavy) {
            if ((e = cudaMemsetAsync(b.heavy_n, 0, sizeof(uint32_t), s)) != cudaSuccess) return e;
            msm_heavy_list_kernel<<<blocks_for(nb, 256), 256, 0, s>>>(nb, b.counts, b.heavy_n, (MsmHeavyItem*)b.heavy_items,
                                                                      b.max_heavy, (uint32_t)pl.seg);
            B200_COUNT_LAUNCH();
        }
        // buckets in order of decreasing size (equal work per warp), then one thread per bucket
        if ((e = cudaMemsetAsync(b.size_hist, 0, 2 * B200_MSM_SIZE_BINS * sizeof(uint32_t), s)) != cudaSuccess) return e;
        msm_size_hist_kernel<<<blocks_for(nb, 256), 256, 0, s>>>(nb, b.counts, b.size_hist);
        msm_size_scan_kernel<<<1, B200_MSM_SIZE_BINS, 0, s>>>(b.size_hist, b.size_hist + B200_MSM_SIZE_BINS);
        msm_size_scatter_kernel<<<blocks_for(nb, 256), 256, 0, s>>>(nb, b.counts, b.size_hist + B200_MSM_SIZE_BINS, b.perm, 0u);
        B200_COUNT_LAUNCH(); B200_COUNT_LAUNCH(); B200_COUNT_LAUNCH();
        if (b.points_ready && (e = cudaStreamWaitEvent(s, b.points_ready, 0)) != cudaSuccess) return e;
        msm_accumulate_kernel<C, G, (IS_G1 ? B200_MSM_ACC_MIN_BLOCKS : 2)><<<blocks_for(nb, 128), 128, 0, s>>>(
            n, pl, (const Aff*)pts, b.offsets, b.counts, b.sorted, b.perm, (Pt*)b.buckets, nb);
        B200_COUNT_LAUNCH();
        if (b.max_heavy) {
            // long runs (more than B200_MSM_SEG points in one bucket): split over extra threads; empty for uniform scalars
            const unsigned hb = blocks_for(b.max_heavy, 128);
            msm_heavy_accumulate_kernel<C, G><<<hb, 128, 0, s>>>(n, pl, (const Aff*)pts, b.offsets, b.counts, b.sorted,
                                                              b.heavy_n, (const MsmHeavyItem*)b.heavy_items,
                                                              (Pt*)b.heavy_partial);
            msm_heavy_merge_kernel<C, G><<<hb, 128, 0, s>>>(b.counts, b.heavy_n, (const MsmHeavyItem*)b.heavy_items,
                                                         (const Pt*)b.heavy_partial, (Pt*)b.buckets, (uint32_t)pl.seg);
            B200_COUNT_LAUNCH(); B200_COUNT_LAUNCH();
        }
        // (Tried in round 2: accumulating the windows in four groups from the top while side streams reduce each finished
        // group and run its share of the Horner chain.  The split accumulation lost more -- 6.0 -> 9.1 ms, the top window's
        // large buckets no longer hide behind the other windows -- than the overlapped tail won)
        MsmPlan pt = pl;                             // plan of the tail (one window when the bucket arrays were folded)
        if constexpr (IS_G1) {
            if (pl.tables && pl.W > 1) {
                msm_fold_kernel<C><<<blocks_for(pl.B, 128), 128, 0, s>>>(pl, (Pt*)b.buckets);
                B200_COUNT_LAUNCH();
                pt.W = 1;
            }
        }
        msm_reduce_kernel<C, G><<<blocks_for((size_t)pt.W * pt.nchunks, 128), 128, 0, s>>>(pt, (const Pt*)b.buckets,
                                                                                        (Pt*)b.chunks);
        B200_COUNT_LAUNCH();
        if (pt.nchunks >= 1024) {
            // two stages: 8 partial sums per window (buffer behind the W window slots), then the windows
            Pt* part = (Pt*)b.windows + pl.W;
            msm_window_sum_kernel<C, G><<<pt.W * 8, 64, 64 * sizeof(Pt), s>>>(pt.nchunks / 8, (const Pt*)b.chunks, part);
            msm_window_sum_kernel<C, G><<<pt.W, 32, 32 * sizeof(Pt), s>>>(8, (const Pt*)part, (Pt*)b.windows);
            B200_COUNT_LAUNCH();
        } else {
            msm_window_sum_kernel<C, G><<<pt.W, 64, 64 * sizeof(Pt), s>>>(pt.nchunks, (const Pt*)b.chunks, (Pt*)b.windows);
        }
        B200_COUNT_LAUNCH();
        if constexpr (IS_G1) msm_final_kernel<C><<<1, 32, 0, s>>>(pt, (const Pt*)b.windows, out, flags);
        else msm_final_g2_kernel<C><<<1, 32, 0, s>>>(pt, (const Pt*)b.windows, out, flags);
        B200_COUNT_LAUNCH();
        return cudaGetLastError();
    }
    static cudaError_t msm(size_t n, const void* pts, const uint8_t* scalars, uint8_t* out, uint32_t flags,
                           const MsmPlan& pl, const MsmBuffers& b, cudaStream_t s) {
        return msm_impl<G1Ops<C>>(n, pts, scalars, out, flags, pl, b, s);
    }
    static cudaError_t msm_g2(size_t n, const void* pts, const uint8_t* scalars, uint8_t* out, uint32_t flags,
                              const MsmPlan& pl, const MsmBuffers& b, cudaStream_t s) {
        return msm_impl<G2Ops<C>>(n, pts, scalars, out, flags, pl, b, s);
    }
    static cudaError_t msm_points_g2(size_t n, const uint8_t* pts, void* out, uint32_t flags, int* err, cudaStream_t s) {
        if (n == 0) return cudaSuccess;
        msm_points_g2_kernel<C><<<blocks_for(n, 128), 128, 0, s>>>(n, pts, (G2Aff<C::N>*)out, flags, err);
        B200_COUNT_LAUNCH();
        return cudaGetLastError();
    }
    static const CurveVTable* table() {
        static const CurveVTable t = {C::FP_BYTES, C::N, C::SCALAR_BITS, sizeof(G1Affine<C::N>), sizeof(G1XYZZ<C::N>),
                                      C::GLV_BITS,
                                      &pairing, &fexp, &g1_mul, &g1_mul2, &g1_sum, &g2_mul, &g2_sum, &gt_op, &lines_row_words, &lines_build, &pairing_fixed, &point_codec, &g1_normalize,
                                      (C::N == 12 && C::BETA == -1) ? &hash_to_g1 : nullptr, &msm_points, &msm_tables, &msm, &msm_points_g2, &msm_g2};
        return &t;
    }
};
#endif

}  // namespace b200
