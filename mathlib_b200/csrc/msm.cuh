// Pippenger G1 multi-scalar multiplication (driver.Curve.MultiScalarMul, reference driver/math.go:170;
// gnark MultiExp call sites bn254.go:232-245, bls12-377.go:229-242, bls12381/bls12-381.go:766-783).
//
// Pipeline (all on device, one stream):
//   1 msm_points_kernel   BYTES -> Montgomery affine points (skipped for MONT input / resident bases)
//   2 msm_digits_kernel   scalar -> (mod r) -> W signed c-bit digits; per-(window,bucket) histogram
//   3 msm_scan_kernel     exclusive scan of the histogram per window
//   4 msm_scatter_kernel  counting-sort point indices by (window, bucket)
//   5 msm_accumulate_kernel  one thread per bucket: XYZZ mixed adds over its sorted run (coalesced index
//                            reads, 2*FpBytes gather per point out of L2/HBM)
//   6 msm_reduce_kernel   sum_b b*B_b per window: chunked running sums, each chunk scaled by its base index
//   7 msm_window_sum_kernel  per window: tree-sum of the chunk results in shared memory
//   8 msm_final_kernel    Horner over the windows (c doublings each), affine normalisation, store
// Resident bases uploaded with B200_BASES_TABLES carry the multiples 2^(c*w) * P_i for every window (W x the memory:
// 1.6 GB for 2^20 BLS12-381 points, small against 180 GB of HBM).  Every window then sums into the SAME bucket index
// space: msm_fold_kernel adds the W bucket arrays, steps 6-8 run on one window and the 240-doubling Horner chain
// disappears.
// The result is a canonical group element, so the order of additions inside a bucket does not matter.
#pragma once
#include <cstdlib>
#include "kernels.cuh"
#include "g2.cuh"

namespace b200 {

#define B200_MSM_SEG 512          // default segment length of a long bucket run (MsmPlan.seg)
struct MsmPlan {
    int c;          // window bits
    int W;          // windows
    int B;          // buckets per window = 2^(c-1)
    int chunk;      // buckets per reduce thread
    int nchunks;    // B / chunk
    int tables;     // 1: points come from a resident window table, tab[w*stride + i] = 2^start(w) * P_i  (no Horner tail)
    unsigned long long stride;
    int narrow;     // the top `narrow` windows are c-1 bits wide (see msm_plan)
    int seg;        // a bucket accumulates at most `seg` points on one thread; longer runs are split (msm_heavy_*)
    int glv;        // 1: scalars are split k = k1 + lambda k2 (130-bit halves) over 2n points P_i, phi(P_i) (see msm_plan_glv)
};

// Window w covers scalar bits [start, start + width).  The W windows tile the scalar exactly: the lower W - narrow ones
// are c bits wide, the top `narrow` ones c - 1 bits.  (Equal c-bit windows would leave a top window of bits mod c bits --
// or only the carry -- whose few buckets collect n / 2^(bits mod c) points each and serialise the accumulation.)
static inline
#if defined(__CUDACC__)
__host__ __device__
#endif
int msm_win_width(const MsmPlan& p, int w) { return w < p.W - p.narrow ? p.c : p.c - 1; }
static inline
#if defined(__CUDACC__)
__host__ __device__
#endif
int msm_win_start(const MsmPlan& p, int w) {
    const int full = p.W - p.narrow;
    return w <= full ? w * p.c : full * p.c + (w - full) * (p.c - 1);
}

static inline MsmPlan msm_plan(size_t n, int scalar_bits, int force_c = 0) {
    int lg = 0;
    while (((size_t)1 << (lg + 1)) <= n) lg++;
    // window size: measured on B200 (tools/msm_sizes_probe.py).  Below ~2^17 points the run time is the latency of the
    // serial tails (reduce / window sums / Horner), which shrink with FEWER windows, so c is larger than the work-optimal
    // lg - 3; from 2^17 on the bucket accumulation dominates and c = 16.
#if defined(B200_DEV_KNOBS)
    static const int bias = getenv("B200_MSM_C_BIAS") ? atoi(getenv("B200_MSM_C_BIAS")) : 0;     // development knob
#else
    const int bias = 0;
#endif
    int c = lg <= 13 ? (lg + 1 > 7 ? lg + 1 : 7) : (lg == 14 ? 15 : (lg == 15 ? 14 : (lg == 16 ? 15 : 16)));
    c += bias;
    if (c < 2) c = 2;
    if (c > 16) c = 16;
    if (force_c) c = force_c;
    MsmPlan p;
    p.c = c;
    p.W = scalar_bits / c + 1;
    // widths in {c, c-1} summing to scalar_bits, at least one narrow window on top: its digit plus the incoming carry
    // is at most 2^(c-1) = B, so the top window needs no carry out
    p.narrow = p.W * c - scalar_bits;
    if (p.narrow > p.W) {                 // cannot happen for c <= 16 and 253..255-bit orders; keep the plan valid anyway
        p.W = (scalar_bits + c - 2) / (c - 1);
        p.narrow = p.W;
    }
    p.B = 1 << (c - 1);
#if defined(B200_DEV_KNOBS)
    static const int chunk_knob = getenv("B200_MSM_CHUNK") ? atoi(getenv("B200_MSM_CHUNK")) : 0;    // development knob
#else
    const int chunk_knob = 0;
#endif
    p.chunk = p.B >= 1024 ? (chunk_knob ? chunk_knob : 16) : (p.B >= 32 ? 8 : 1);
    p.nchunks = p.B / p.chunk;
    p.tables = 0;
    p.stride = 0;
    p.glv = 0;
    p.seg = B200_MSM_SEG;
    return p;
}

// GLV plan (BLS12 curves, one-shot MSMs): every scalar is split exactly, k = k1 + lambda k2 with
// 0 <= k1 < lambda and k2 <= r / lambda, both below 2^glv_bits (128 on BLS12-381, 127 on BLS12-377; g1.cuh glv_split plus up
// to three corrections in msm_digits_kernel); [lambda](x, y) = (beta x, y) costs one Fp product per point when the points
// are converted.  The MSM then runs over 2n points and glv_bits-bit scalars: half the windows for the same number of
// bucket additions, and the serial Horner tail is 112 doublings instead of 240.  At 2^20 points and up: eight 16-bit
// windows (c = 17, all narrow: signed digits in windows 0..6, the top window unsigned over 2^16 buckets -- an exact split
// matters here: with the 130-bit slack of the approximate split the top window's digits would use a sixteenth of its
// buckets and those would be several times fuller than the rest).
static inline MsmPlan msm_plan_glv(size_t n, int glv_bits) {
    int lg = 0;
    while (((size_t)1 << (lg + 1)) <= 2 * n) lg++;
    const int c = lg >= 20 ? 17 : (lg >= 17 ? 15 : (lg >= 15 ? 14 : 0));      // 0: msm_plan's choice for small inputs
    MsmPlan p = msm_plan(2 * n, glv_bits, c);
    p.glv = 1;
    // signed digits leave half of a narrow window's buckets empty, so half of the reduce threads have nothing to do:
    // chunks of 8 buckets (not 16) keep two warps per sub-partition busy (2^20: 8.84 -> 8.65 ms, 2^17: 2.80 -> 2.59 ms;
    // 4: 9.25 / 2.66 ms, 32: 9.01 / 3.34 ms)
    if (p.B >= 1024) { p.chunk = 8; p.nchunks = p.B / p.chunk; }
    return p;
}
#ifndef B200_MSM_GLV_MIN
#define B200_MSM_GLV_MIN ((size_t)1)      // every size gains: half the windows means half the serial tail (2^4: 1.96 -> 1.28 ms, 2^13: 2.41 -> 1.79 ms)
#endif

#if defined(__CUDACC__)

// glv != 0: also out[n + i] = phi(P_i) = (beta x, y)  (infinity (0, 0) maps to itself)
template <class C>
__global__ void msm_points_kernel(size_t n, const uint8_t* pts, G1Affine<C::N>* out, uint32_t flags, int* err, int glv) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int e = 0;
    G1Affine<C::N> a;
    Codec<C>::g1_load(a.x, a.y, pts + i * Codec<C>::g1_size(), flags & FLAG_IN_MONT, &e);
    if (e) atomicExch(err, 1);
    out[i] = a;
    if (glv) {
        Fp<C::N> beta;
        const uint32_t* bw = C::K().glv_beta;
        for (int k = 0; k < C::N; k++) beta.l[k] = bw[k];
        FpOps<C>::mulx(a.x, a.x, beta);
        out[n + i] = a;
    }
}

// k (8 LE words) mod r by repeated conditional subtraction (k < 2^256 < 14 r for all three curves)
template <class C>
__device__ __forceinline__ void scalar_reduce(uint32_t* k) {
    const uint32_t* r = C::K().order;
    for (int it = 0; it < 16; it++) {
        uint32_t t[8];
        t[0] = sub_cc(k[0], r[0]);
#pragma unroll
        for (int i = 1; i < 8; i++) t[i] = subc_cc(k[i], r[i]);
        uint32_t borrow = subc(0, 0);
        if (borrow) break;
#pragma unroll
        for (int i = 0; i < 8; i++) k[i] = t[i];
    }
}

// digits[w*stride + idx] = signed digit of the (<= 256-bit, zero-extended to 9 words) scalar k in window w, packed as
// (|d| << 1) | sign ; 0 = skip
__device__ __forceinline__ void msm_emit_digits(const uint32_t* k, size_t idx, size_t stride, const MsmPlan& pl, uint32_t* digits,
                                                uint32_t* counts) {
    uint32_t carry = 0;
    for (int w = 0; w < pl.W; w++) {
        const int bit = msm_win_start(pl, w), cw = msm_win_width(pl, w);
        uint32_t lo = k[bit >> 5] >> (bit & 31);
        if ((bit & 31) + cw > 32 && (bit >> 5) + 1 < 9) lo |= k[(bit >> 5) + 1] << (32 - (bit & 31));
        uint32_t d = (bit < 256 ? (lo & ((1u << cw) - 1)) : 0) + carry;
        uint32_t neg = 0;
        // signed digits: fold d > 2^(cw-1) to d - 2^cw with a carry into the next window; never on the top window
        if (w + 1 < pl.W && d > (1u << (cw - 1))) { d = (1u << cw) - d; neg = 1; carry = 1; } else carry = 0;
        uint32_t packed = d ? ((d << 1) | neg) : 0;
        digits[(size_t)w * stride + idx] = packed;
        if (d) atomicAdd(&counts[(size_t)w * pl.B + (d - 1)], 1u);
    }
}
template <class C>
__global__ void msm_digits_kernel(size_t n, const uint8_t* scalars, MsmPlan pl, uint32_t* digits, uint32_t* counts) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k[9];
    Codec<C>::scalar_load(k, scalars + i * 32);
    scalar_reduce<C>(k);
    k[8] = 0;
    if (!pl.glv) {
        msm_emit_digits(k, i, n, pl, digits, counts);
        return;
    }
    // GLV: point i takes k1, point n + i (= phi(P_i)) takes k2
    uint32_t h[9];
    uint32_t k1[5], k2[5];
    G1Ops<C>::glv_split(k1, k2, k);
    // glv_split under-estimates k2 by at most 3: make the split exact (k1 < lambda) so that both halves fit GLV_BITS
    {
        const uint32_t* lam = C::K().glv_lambda;
        for (int it = 0; it < 3; it++) {
            uint32_t t[5];
            t[0] = sub_cc(k1[0], lam[0]);
            t[1] = subc_cc(k1[1], lam[1]);
            t[2] = subc_cc(k1[2], lam[2]);
            t[3] = subc_cc(k1[3], lam[3]);
            t[4] = subc_cc(k1[4], 0);
            const uint32_t borrow = subc(0, 0);
            if (borrow) break;
            for (int j = 0; j < 5; j++) k1[j] = t[j];
            k2[0] = add_cc(k2[0], 1);
            for (int j = 1; j < 4; j++) k2[j] = addc_cc(k2[j], 0);
            k2[4] = addc(k2[4], 0);
        }
    }
    for (int j = 0; j < 9; j++) h[j] = j < 5 ? k1[j] : 0u;
    msm_emit_digits(h, i, 2 * n, pl, digits, counts);
    for (int j = 0; j < 5; j++) h[j] = k2[j];
    msm_emit_digits(h, n + i, 2 * n, pl, digits, counts);
}

// per window exclusive scan of counts[w*B .. w*B+B) -> offsets (relative to the window), one block per window
static __global__ void msm_scan_kernel(MsmPlan pl, const uint32_t* counts, uint32_t* offsets) {
    extern __shared__ uint32_t sh[];
    int w = blockIdx.x;
    const uint32_t* c = counts + (size_t)w * pl.B;
    uint32_t* o = offsets + (size_t)w * pl.B;
    int per = (pl.B + blockDim.x - 1) / blockDim.x;
    int lo = threadIdx.x * per, hi = min(lo + per, pl.B);
    uint32_t s = 0;
    for (int j = lo; j < hi; j++) s += c[j];
    sh[threadIdx.x] = s;
    __syncthreads();
    // simple Hillis-Steele inclusive scan over blockDim.x partial sums
    for (int d = 1; d < blockDim.x; d <<= 1) {
        uint32_t v = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
        __syncthreads();
        sh[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t base = sh[threadIdx.x] - s;
    for (int j = lo; j < hi; j++) { o[j] = base; base += c[j]; }
}

// sorted[w*n + offsets[w,b] + slot] = (i << 1) | sign ; cursor starts as a copy of offsets
static __global__ void msm_scatter_kernel(size_t n, MsmPlan pl, const uint32_t* digits, uint32_t* cursor, uint32_t* sorted) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int w = 0; w < pl.W; w++) {
        uint32_t p = digits[(size_t)w * n + i];
        if (!p) continue;
        uint32_t b = (p >> 1) - 1;
        uint32_t pos = atomicAdd(&cursor[(size_t)w * pl.B + b], 1u);
        sorted[(size_t)w * n + pos] = ((uint32_t)i << 1) | (p & 1);
    }
}

// Load balancing: buckets are handed to threads in order of decreasing size (counting sort of the bucket ids by
// their point count, sizes clamped to 1023), so the 32 threads of a warp run the same number of mixed additions.
#define B200_MSM_SIZE_BINS 1024
// Empty buckets are common (signed digits leave half of a narrow window's buckets unused; small MSMs), and they would all
// hit bin 0: their atomics are aggregated per warp (one atomicAdd per warp instead of up to 32).
static __global__ void msm_size_hist_kernel(size_t nb, const uint32_t* counts, uint32_t* hist) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < nb;
    const uint32_t c = live ? counts[t] : 1u;
    const unsigned zero = __ballot_sync(0xffffffffu, live && c == 0);
    if (zero && (threadIdx.x & 31) == (unsigned)(__ffs(zero) - 1)) atomicAdd(&hist[0], (uint32_t)__popc(zero));
    if (live && c) atomicAdd(&hist[c < B200_MSM_SIZE_BINS ? c : B200_MSM_SIZE_BINS - 1], 1u);
}
// exclusive scan of hist in DESCENDING size order -> start[]; single block of B200_MSM_SIZE_BINS threads
static __global__ void msm_size_scan_kernel(const uint32_t* hist, uint32_t* start) {
    __shared__ uint32_t sh[B200_MSM_SIZE_BINS];
    int i = threadIdx.x;                         // rank 0 = largest size
    uint32_t v = hist[B200_MSM_SIZE_BINS - 1 - i];
    sh[i] = v;
    __syncthreads();
    for (int d = 1; d < B200_MSM_SIZE_BINS; d <<= 1) {
        uint32_t x = i >= d ? sh[i - d] : 0;
        __syncthreads();
        sh[i] += x;
        __syncthreads();
    }
    start[B200_MSM_SIZE_BINS - 1 - i] = sh[i] - v;
}
static __global__ void msm_size_scatter_kernel(size_t nb, const uint32_t* counts, uint32_t* start, uint32_t* perm, uint32_t base) {
    // counts / perm already point at this launch's range of buckets; the ids written are global (base + local index)
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < nb;
    const uint32_t c = live ? counts[t] : 1u;
    const unsigned lane = threadIdx.x & 31;
    const unsigned zero = __ballot_sync(0xffffffffu, live && c == 0);
    uint32_t zbase = 0;
    if (zero) {
        const int leader = __ffs(zero) - 1;
        if ((int)lane == leader) zbase = atomicAdd(&start[0], (uint32_t)__popc(zero));
        zbase = __shfl_sync(0xffffffffu, zbase, leader);
    }
    if (!live) return;
    uint32_t pos;
    if (c == 0) pos = zbase + (uint32_t)__popc(zero & ((1u << lane) - 1));
    else pos = atomicAdd(&start[c < B200_MSM_SIZE_BINS ? c : B200_MSM_SIZE_BINS - 1], 1u);
    perm[pos] = base + (uint32_t)t;
}

// Long runs (skewed scalars: many equal digits) would serialise on one thread.  A bucket accumulates at most B200_MSM_SEG
// points in msm_accumulate_kernel; every further segment of B200_MSM_SEG points becomes an item (bucket, segment) of a
// list built on the device, is summed by its own thread, and the partial sums are added to the bucket afterwards.  With
// uniform scalars the list is empty and the two extra kernels exit at once.
struct MsmHeavyItem { uint32_t bucket, seg; };
static __global__ void msm_heavy_list_kernel(size_t nb, const uint32_t* counts, uint32_t* heavy_n, MsmHeavyItem* items,
                                             uint32_t max_items, uint32_t seg) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nb) return;
    const uint32_t cnt = counts[t];
    if (cnt <= seg) return;
    const uint32_t k = (cnt - 1) / seg;                                // extra segments 1..k
    const uint32_t pos = atomicAdd(heavy_n, k);
    for (uint32_t s = 1; s <= k && pos + s - 1 < max_items; s++) items[pos + s - 1] = {(uint32_t)t, s};
}
// The bucket kernels below are generic in the group: G = G1Ops<C> (driver.Curve.MultiScalarMul) or G2Ops<C> (the G2 MSM,
// SURVEY 8(f) row 3) -- same digits, sort and schedule, XYZZ formulas over Fp or Fp2.
template <class C, class G = G1Ops<C>>
__global__ void __launch_bounds__(128, 2)
msm_heavy_accumulate_kernel(size_t n, MsmPlan pl, const typename G::Aff* pts, const uint32_t* offsets, const uint32_t* counts,
                            const uint32_t* sorted, const uint32_t* heavy_n, const MsmHeavyItem* items, typename G::Pt* partial) {
    size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= *heavy_n) return;
    const MsmHeavyItem it = items[id];
    const size_t t = it.bucket, w = t / pl.B;
    if (pl.tables) pts += w * pl.stride;
    const uint32_t lo = it.seg * (uint32_t)pl.seg;
    uint32_t hi = lo + (uint32_t)pl.seg;
    if (hi > counts[t]) hi = counts[t];
    const uint32_t* run = sorted + w * n + offsets[t];
    typename G::Pt acc;
    G::set_inf(acc);
    for (uint32_t j = lo; j < hi; j++) {
        const uint32_t e = run[j];
        typename G::Aff a = pts[e >> 1];
        if (e & 1) G::neg_y(a);
        G::madd(acc, a);
    }
    partial[id] = acc;
}
// one thread per heavy bucket (the thread of its first item): the items of a bucket are contiguous in the list
template <class C, class G = G1Ops<C>>
__global__ void __launch_bounds__(128, 2)
msm_heavy_merge_kernel(const uint32_t* counts, const uint32_t* heavy_n, const MsmHeavyItem* items, const typename G::Pt* partial,
                       typename G::Pt* buckets, uint32_t seg) {
    size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= *heavy_n) return;
    const MsmHeavyItem it = items[id];
    if (it.seg != 1) return;
    const uint32_t k = (counts[it.bucket] - 1) / seg;
    typename G::Pt acc = buckets[it.bucket];
    for (uint32_t s = 0; s < k; s++) {
        typename G::Pt v = partial[id + s];
        G::add(acc, v);
    }
    buckets[it.bucket] = acc;
}

#ifndef B200_MSM_ACC_MIN_BLOCKS
#define B200_MSM_ACC_MIN_BLOCKS 4      // measured at 2^20 points: 2 blocks (188 regs) 9.64 ms, 3 (168) 9.25 ms, 4 (128, a few spills) 9.17 ms
#endif
template <class C, class G = G1Ops<C>, int MINB = B200_MSM_ACC_MIN_BLOCKS>
__global__ void __launch_bounds__(128, MINB)
msm_accumulate_kernel(size_t n, MsmPlan pl, const typename G::Aff* pts, const uint32_t* offsets, const uint32_t* counts,
                      const uint32_t* sorted, const uint32_t* perm, typename G::Pt* buckets, size_t nb) {
    // perm[0..nb): the (window, bucket) ids this launch works on, largest buckets first
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= nb) return;
    const size_t t = perm[tid];
    size_t w = t / pl.B;
    if (pl.tables) pts += w * pl.stride;
    const uint32_t* run = sorted + w * n + offsets[t];
    uint32_t cnt = counts[t];
    if (cnt > (uint32_t)pl.seg) cnt = (uint32_t)pl.seg;  // the rest of a long run is split over msm_heavy_* threads
    typename G::Pt acc;
    G::set_inf(acc);
    // software prefetch: the gather of point j+1 (index read + 2*FpBytes random read out of L2/HBM) is issued before
    // the mixed addition of point j
    typename G::Aff nxt;
    uint32_t e_nxt = 0;
    if (cnt) { e_nxt = run[0]; nxt = pts[e_nxt >> 1]; }
    for (uint32_t j = 0; j < cnt; j++) {
        typename G::Aff a = nxt;
        const uint32_t e = e_nxt;
        if (j + 1 < cnt) { e_nxt = run[j + 1]; nxt = pts[e_nxt >> 1]; }
        if (e & 1) G::neg_y(a);
        G::madd(acc, a);
    }
    buckets[t] = acc;
}

// window tables: tab[w*stride + i] = 2^(c*w) * tab[i], affine.  One thread per point walks the windows (c doublings and
// one inversion each); runs once per upload.
template <class C>
__global__ void __launch_bounds__(128)
msm_tables_kernel(size_t n, MsmPlan pl, size_t stride, G1Affine<C::N>* tab) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    typedef G1Ops<C> G;
    typename G::Aff a = tab[i];
    for (int w = 1; w < pl.W; w++) {
        typename G::Pt p;
        G::dbl_affine(p, a);
        for (int k = 1; k < msm_win_width(pl, w - 1); k++) G::dbl(p);
        G::to_affine(a, p);
        tab[(size_t)w * stride + i] = a;
    }
}

// window tables: buckets[0][b] += sum_{w>=1} buckets[w][b]
template <class C>
__global__ void __launch_bounds__(128)
msm_fold_kernel(MsmPlan pl, G1XYZZ<C::N>* buckets) {
    size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= (size_t)pl.B) return;
    typedef G1Ops<C> G;
    typename G::Pt acc = buckets[b];
    for (int w = 1; w < pl.W; w++) {
        typename G::Pt v = buckets[(size_t)w * pl.B + b];
        G::add(acc, v);
    }
    buckets[b] = acc;
}

// thread (w, chunk t): G = sum_{j=1..S} (t*S + j) * B[w][t*S + j - 1]
template <class C, class G = G1Ops<C>>
__global__ void __launch_bounds__(128)
msm_reduce_kernel(MsmPlan pl, const typename G::Pt* buckets, typename G::Pt* chunk_out) {
    size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (size_t)pl.W * pl.nchunks) return;
    size_t w = id / pl.nchunks, t = id % pl.nchunks;
    const typename G::Pt* b = buckets + w * pl.B + t * pl.chunk;
    typename G::Pt run, acc;
    G::set_inf(run);
    G::set_inf(acc);
    for (int j = pl.chunk - 1; j >= 0; j--) {
        typename G::Pt v = b[j];
        G::add(run, v);
        G::add(acc, run);
    }
    // acc = sum (j+1) B_j ; run = sum B_j ; add (t*S) * run
    uint32_t m = (uint32_t)(t * pl.chunk);
    if (m) {
        typename G::Pt s;
        G::set_inf(s);
        for (int bit = 31 - __clz(m); bit >= 0; bit--) {
            G::dbl(s);
            if ((m >> bit) & 1) G::add(s, run);
        }
        G::add(acc, s);
    }
    chunk_out[id] = acc;
}

// segment sums: out[seg] = sum of in[seg*count .. seg*count + count), one block per segment (shared-memory tree).
// The per-window sum of the chunk results runs in two stages (W*8 segments, then W) so that it spreads over 128 SMs
// instead of 16.
template <class C, class G = G1Ops<C>>
__global__ void msm_window_sum_kernel(int count, const typename G::Pt* in, typename G::Pt* out) {
    extern __shared__ uint32_t shraw[];
    typename G::Pt* sh = reinterpret_cast<typename G::Pt*>(shraw);
    const size_t seg = blockIdx.x;
    typename G::Pt acc;
    G::set_inf(acc);
    for (int j = threadIdx.x; j < count; j += blockDim.x) {
        typename G::Pt v = in[seg * count + j];
        G::add(acc, v);
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int d = blockDim.x / 2; d > 0; d >>= 1) {
        if (threadIdx.x < d) {
            typename G::Pt a = sh[threadIdx.x], b = sh[threadIdx.x + d];
            G::add(a, b);
            sh[threadIdx.x] = a;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[seg] = sh[0];
}

// Horner over the windows, then affine normalisation + store.  One thread: this is a pure dependency chain of
// c*(W-1) doublings, so it is written for latency -- Jacobian doubling (2M+5S) with the Fp products inlined so the
// independent ones overlap in the pipe, instead of the out-of-line XYZZ formulas.
template <class C>
struct JacOps {
    typedef FpOps<C> F;
    typedef Fp<C::N> E;
    struct Pt { E x, y, z; };          // z == 0 <=> infinity
    static __device__ __forceinline__ void dbl(Pt& p) {          // dbl-2009-l, a = 0
        E A, B, Cc, D, Ev, Fv, t;
        F::sqr(A, p.x);
        F::sqr(B, p.y);
        F::mul(t, p.y, p.z);
        F::sqr(Cc, B);
        F::add(D, p.x, B);
        F::sqr(D, D);
        F::sub(D, D, A); F::sub(D, D, Cc); F::dbl(D, D);
        F::dbl(Ev, A); F::add(Ev, Ev, A);
        F::sqr(Fv, Ev);
        F::dbl(p.z, t);
        F::sub(p.x, Fv, D); F::sub(p.x, p.x, D);
        F::sub(t, D, p.x);
        F::mul(t, Ev, t);
        F::dbl(Cc, Cc); F::dbl(Cc, Cc); F::dbl(Cc, Cc);
        F::sub(p.y, t, Cc);
    }
};

// One warp: acc = 2^tail * sum_{w in [w_lo, w_hi)} 2^(start(w) - start(w_lo)) windows[w]  (Horner from the top window down, then
// `tail` more doublings).  The chain is serial in the doublings, but inside a doubling three field products are independent
// at each of the first two levels: lanes 0..2 take one each (every lane keeps a full copy of the point and repeats the
// cheap additions), results are exchanged through shared memory.  4 product latencies per doubling instead of 8.
template <class C>
__device__ __forceinline__ void msm_horner_warp(const MsmPlan& pl, int w_lo, int w_hi, int tail, const G1XYZZ<C::N>* windows,
                                                G1XYZZ<C::N>& acc, Fp<C::N>* sh) {
    typedef G1Ops<C> G;
    typedef FpOps<C> F;
    typedef JacOps<C> J;
    typedef typename F::E E;
    const int lane = threadIdx.x & 31;
    const int sel = lane < 2 ? lane : 2;
    // r0, r1, r2 = a0*b0, a1*b1, a2*b2, one product per lane
    auto par3 = [&](E& r0, E& r1, E& r2, const E& a0, const E& b0, const E& a1, const E& b1, const E& a2, const E& b2) {
        E a = sel == 0 ? a0 : (sel == 1 ? a1 : a2);
        E b = sel == 0 ? b0 : (sel == 1 ? b1 : b2);
        E r;
        F::mul(r, a, b);
        if (lane < 3) sh[lane] = r;
        __syncwarp();
        r0 = sh[0]; r1 = sh[1]; r2 = sh[2];
        __syncwarp();
    };
    // acc <- 2^cnt acc through Jacobian coordinates
    auto doublings = [&](int cnt) {
        if (G::is_inf(acc) || cnt <= 0) return;
        // XYZZ (X, Y, ZZ, ZZZ) -> Jacobian with Z' = ZZ*ZZZ:  x = X/ZZ = (X*ZZ*ZZZ^2)/Z'^2, y = Y/ZZZ = (Y*ZZ^3*ZZZ^2)/Z'^3
        typename J::Pt j;
        E zz2, zzz2, t, u;
        par3(zzz2, zz2, j.z, acc.zzz, acc.zzz, acc.zz, acc.zz, acc.zz, acc.zzz);
        par3(t, u, zz2, acc.x, acc.zz, acc.y, zz2, zz2, zz2);           // t = X*ZZ, u = Y*ZZ^2 (third product unused)
        F::mul(u, u, acc.zz);
        par3(j.x, j.y, zz2, t, zzz2, u, zzz2, t, t);
        for (int k = 0; k < cnt; k++) {                                  // dbl-2009-l, a = 0
            E A, B, T, Cc, D, Fv, Ev, s;
            par3(A, B, T, j.x, j.x, j.y, j.y, j.y, j.z);
            F::add(s, j.x, B);
            F::dbl(Ev, A); F::add(Ev, Ev, A);
            par3(Cc, D, Fv, B, B, s, s, Ev, Ev);
            F::sub(D, D, A); F::sub(D, D, Cc); F::dbl(D, D);
            F::dbl(j.z, T);
            F::sub(j.x, Fv, D); F::sub(j.x, j.x, D);
            F::sub(s, D, j.x);
            F::mul(s, Ev, s);
            F::dbl(Cc, Cc); F::dbl(Cc, Cc); F::dbl(Cc, Cc);
            F::sub(j.y, s, Cc);
        }
        // back to XYZZ: ZZ = Z^2, ZZZ = Z^3
        acc.x = j.x; acc.y = j.y;
        F::sqr(acc.zz, j.z);
        F::mul(acc.zzz, acc.zz, j.z);
    };
    G::set_inf(acc);
    for (int w = w_hi - 1; w >= w_lo; w--) {
        if (w + 1 < w_hi) doublings(msm_win_width(pl, w));
        typename G::Pt v = windows[w];
        // acc += v (add-2008-s) with the independent products spread over the three lanes: 5 product latencies, not 14.
        // Infinity operands and equal / opposite points (warp-uniform conditions) take the complete serial adder.
        bool generic = !G::is_inf(acc) && !G::is_inf(v);
        E U1, U2, S1, S2, PP, ZZ12, Pp;
        if (generic) {
            par3(U1, U2, S1, acc.x, v.zz, v.x, acc.zz, acc.y, v.zzz);
            F::sub(Pp, U2, U1);
            generic = !F::is_zero(Pp);
        }
        if (!generic) {
            G::add(acc, v);
            continue;
        }
        E PPP, Q, ZZZ12, R, R2, t, u;
        par3(S2, PP, ZZ12, v.y, acc.zzz, Pp, Pp, acc.zz, v.zz);
        par3(PPP, Q, ZZZ12, Pp, PP, U1, PP, acc.zzz, v.zzz);
        F::sub(R, S2, S1);
        par3(R2, acc.zz, acc.zzz, R, R, ZZ12, PP, ZZZ12, PPP);
        F::sub(t, R2, PPP); F::sub(t, t, Q); F::sub(t, t, Q);            // X3
        F::sub(Q, Q, t);
        par3(u, S1, R2, R, Q, S1, PPP, R, R);                             // R*(Q - X3), S1*PPP (third product unused)
        F::sub(acc.y, u, S1);
        acc.x = t;
    }
    doublings(tail);
}

// Horner over all windows, affine normalisation + store (one warp)
template <class C>
__global__ void msm_final_kernel(MsmPlan pl, const G1XYZZ<C::N>* windows, uint8_t* out, uint32_t flags) {
    if (blockIdx.x != 0) return;
    typedef G1Ops<C> G;
    __shared__ Fp<C::N> sh[3];
    typename G::Pt acc;
    msm_horner_warp<C>(pl, 0, pl.W, 0, windows, acc, sh);
    typename G::Aff r;
    G::to_affine(r, acc);
    if ((threadIdx.x & 31) == 0) Codec<C>::g1_store(out, r.x, r.y, flags & FLAG_OUT_MONT);
}

// ---- G2 MSM: the same pipeline over E'(Fp2).  Points: reference G2.Bytes() encodings (or Montgomery
// slabs) -> Montgomery affine; tail: Horner over the windows in XYZZ on one thread (not latency-tuned like the G1 tail).
template <class C>
__global__ void msm_points_g2_kernel(size_t n, const uint8_t* pts, G2Aff<C::N>* out, uint32_t flags, int* err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int e = 0;
    G2Aff<C::N> a;
    Codec<C>::g2_load(a, pts + i * Codec<C>::g2_size(), flags & FLAG_IN_MONT, &e);
    if (e) atomicExch(err, 1);
    out[i] = a;
}
template <class C>
__global__ void msm_final_g2_kernel(MsmPlan pl, const G2XYZZ<C::N>* windows, uint8_t* out, uint32_t flags) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    typedef G2Ops<C> G;
    typename G::Pt acc;
    G::set_inf(acc);
    for (int w = pl.W - 1; w >= 0; w--) {
        if (!G::is_inf(acc))
            for (int k = 0; k < msm_win_width(pl, w); k++) G::dbl(acc);
        typename G::Pt v = windows[w];
        G::add(acc, v);
    }
    typename G::Aff r;
    G::to_affine(r, acc);
    G2Codec<C>::g2_store(out, r, flags & FLAG_OUT_MONT);
}

#endif  // __CUDACC__
}  // namespace b200
