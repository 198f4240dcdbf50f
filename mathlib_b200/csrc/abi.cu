// C ABI of libb200math (include/b200.h): argument checking, device workspaces, host<->device staging,
// index-split over the GPUs of one box, and the tiny MSM partial-sum combine.  All arithmetic happens in
// the sm_100a kernels reached through the per-curve tables (launch.cuh); there is no CPU fallback.
#include "../../include/b200.h"
#include "launch.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <map>

namespace b200 {
std::atomic<uint64_t> g_launch_count{0};
}
using namespace b200;

namespace {

thread_local std::string t_err;
thread_local int t_device = -1;          // -1: not pinned by b200_set_device
thread_local cudaStream_t t_stream = nullptr;

std::mutex g_mu;
std::atomic<bool> g_inited{false};
std::vector<int> g_devices;          // written under g_mu by b200_init; readers take a snapshot with devices()
std::vector<int> devices() {
    std::lock_guard<std::mutex> lk(g_mu);
    return g_devices;
}

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_err = buf;
    return code;
}
#define CU(x)                                                                                   \
    do {                                                                                        \
        cudaError_t _e = (x);                                                                   \
        if (_e != cudaSuccess) return fail(B200_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(_e)); \
    } while (0)

struct CurveInfo {
    const CurveVTable* vt;
    bool kilic;     // Pairing includes FExp; FExp is the identity
};
bool curve_info(int curve, CurveInfo* ci) {
    switch (curve) {
        case B200_BN254: *ci = {vtable_bn254(), false}; return true;
        case B200_BLS12_381: case B200_BLS12_381_BBS: *ci = {vtable_bls381(), true}; return true;
        case B200_BLS12_381_GURVY: case B200_BLS12_381_BBS_GURVY: *ci = {vtable_bls381(), false}; return true;
        case B200_BLS12_377_GURVY: *ci = {vtable_bls377(), false}; return true;
    }
    return false;
}

// ---------------------------------------------------------------------------------------------
// per-device workspaces (grow-only slab + stream + error flag), pooled so concurrent callers never share one
// ---------------------------------------------------------------------------------------------
struct Workspace {
    int dev = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;    // second stream: the MSM's point upload + conversion overlap the digit / sort kernels
    cudaEvent_t ev_start = nullptr, ev_points = nullptr;
    uint8_t* buf = nullptr;
    size_t cap = 0;
    int* d_err = nullptr;
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        if (buf) { cudaFree(buf); buf = nullptr; cap = 0; }
        size_t want = bytes + bytes / 4 + (1 << 20);
        cudaError_t e = cudaMalloc(&buf, want);
        if (e != cudaSuccess) return fail(B200_ERR_CUDA, "cudaMalloc(%zu): %s", want, cudaGetErrorString(e));
        cap = want;
        return 0;
    }
};
std::map<int, std::vector<Workspace*>> g_free_ws;

Workspace* ws_acquire(int dev) {
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto& v = g_free_ws[dev];
        if (!v.empty()) { Workspace* w = v.back(); v.pop_back(); return w; }
    }
    if (cudaSetDevice(dev) != cudaSuccess) return nullptr;
    Workspace* w = new Workspace();
    w->dev = dev;
    if (cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking) != cudaSuccess) { delete w; return nullptr; }
    if (cudaStreamCreateWithFlags(&w->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&w->ev_start, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&w->ev_points, cudaEventDisableTiming) != cudaSuccess) { delete w; return nullptr; }
    if (cudaMalloc(&w->d_err, sizeof(int)) != cudaSuccess) { delete w; return nullptr; }
    return w;
}
void ws_release(Workspace* w) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_free_ws[w->dev].push_back(w);
}
struct WsGuard {
    Workspace* w;
    explicit WsGuard(int dev) : w(ws_acquire(dev)) {}
    // an error return may leave copies or kernels in flight on the workspace's stream: drain it before the slab goes
    // back to the pool (free when the stream is already idle, which is the normal return path)
    ~WsGuard() {
        if (!w) return;
        cudaSetDevice(w->dev);
        cudaStreamSynchronize(w->stream);
        cudaStreamSynchronize(w->copy_stream);
        ws_release(w);
    }
};

size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

int ensure_init() {
    if (g_inited.load(std::memory_order_acquire)) return 0;
    return b200_init(0);
}
int current_device() { return t_device >= 0 ? t_device : devices()[0]; }

// devices a host-buffer batch call is split over
std::vector<int> split_devices(size_t n, size_t min_per_dev) {
    std::vector<int> all = devices(), d;
    if (t_device >= 0 || all.size() == 1 || n < 2 * min_per_dev) { d.push_back(current_device()); return d; }
    size_t k = all.size();
    while (k > 1 && n / k < min_per_dev) k--;
    d.assign(all.begin(), all.begin() + k);
    return d;
}

// run fn(dev, lo, hi) over contiguous index ranges, one host thread per device
template <class Fn>
int run_split(const std::vector<int>& devs, size_t n, Fn fn) {
    if (devs.size() == 1) return fn(devs[0], (size_t)0, n);
    std::vector<std::thread> th;
    std::vector<int> rc(devs.size(), 0);
    std::vector<std::string> msg(devs.size());
    for (size_t k = 0; k < devs.size(); k++) {
        size_t lo = n * k / devs.size(), hi = n * (k + 1) / devs.size();
        th.emplace_back([&, k, lo, hi]() {
            rc[k] = fn(devs[k], lo, hi);
            if (rc[k]) msg[k] = t_err;
        });
    }
    for (auto& t : th) t.join();
    for (size_t k = 0; k < devs.size(); k++)
        if (rc[k]) { t_err = msg[k]; return rc[k]; }
    return 0;
}

struct Piece { const void* host; size_t elem; uint8_t* dev; };

// stage inputs, run, fetch output for the index range [lo,hi) on one device
template <class LaunchFn>
int staged_call(int dev, size_t lo, size_t hi, std::vector<Piece> ins, void* out_host, size_t out_elem,
                LaunchFn launch) {
    size_t m = hi - lo;
    if (m == 0) return 0;
    WsGuard g(dev);
    if (!g.w) return fail(B200_ERR_CUDA, "cannot create workspace on device %d: %s", dev,
                          cudaGetErrorString(cudaGetLastError()));
    Workspace& w = *g.w;
    CU(cudaSetDevice(dev));
    size_t total = 0;
    for (auto& p : ins) total += align_up(p.elem * m);
    total += align_up(out_elem * m);
    if (int rc = w.reserve(total)) return rc;
    size_t off = 0;
    for (auto& p : ins) {
        p.dev = w.buf + off;
        CU(cudaMemcpyAsync(p.dev, (const uint8_t*)p.host + lo * p.elem, p.elem * m, cudaMemcpyHostToDevice, w.stream));
        off += align_up(p.elem * m);
    }
    uint8_t* d_out = w.buf + off;
    CU(cudaMemsetAsync(w.d_err, 0, sizeof(int), w.stream));
    cudaError_t e = launch(m, ins, d_out, w.d_err, w.stream);
    if (e != cudaSuccess) return fail(B200_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(e));
    int h_err = 0;
    CU(cudaMemcpyAsync((uint8_t*)out_host + lo * out_elem, d_out, out_elem * m, cudaMemcpyDeviceToHost, w.stream));
    CU(cudaMemcpyAsync(&h_err, w.d_err, sizeof(int), cudaMemcpyDeviceToHost, w.stream));
    CU(cudaStreamSynchronize(w.stream));
    if (h_err) return fail(B200_ERR_ENCODING, "input is not a canonical element encoding");
    return 0;
}

// Error flag of the B200_DEVICE_PTRS calls of one device.  Those calls are asynchronous, so a rejected input cannot
// fail them: the kernels write a defined output for the offending item (zero element / verdict 0) and raise this flag;
// b200_take_error() reads and clears it.
int* device_err_flag(int dev) {
    static std::map<int, int*> flags;
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = flags.find(dev);
    if (it != flags.end()) return it->second;
    int* p = nullptr;
    cudaSetDevice(dev);
    cudaMalloc(&p, sizeof(int));
    cudaMemset(p, 0, sizeof(int));
    flags[dev] = p;
    return p;
}

uint32_t kernel_flags(uint32_t flags) { return flags & (B200_FEXP | B200_IN_MONT | B200_OUT_MONT | B200_OUT_UNITY_ONLY); }

struct Bases {
    int curve; int dev; size_t n; void* pts;
    int table_c; int table_w;          // window tables present when table_w > 0 (pts holds table_w * n points)
};
// plan of an MSM of n scalars against resident bases: the window tables are used when their window size is the one
// this n would pick anyway (much smaller calls fall back to the plain points, which are table row 0)
MsmPlan resident_plan(const CurveVTable* vt, const Bases* bs, size_t n) {
    MsmPlan pl = msm_plan(n ? n : 1, vt->scalar_bits);
    if (bs && bs->table_w > 0 && pl.c == bs->table_c && pl.W == bs->table_w) {
        pl.tables = 1;
        pl.stride = bs->n;
    }
    return pl;
}
std::map<uint64_t, Bases> g_bases;
struct Lines {                          // fixed-Q line tables
    int curve; int dev; size_t n_q; uint32_t* lines; uint8_t* qinf;
};
std::map<uint64_t, Lines> g_lines;
uint64_t g_next_handle = 1;

// carve the MSM workspace; returns bytes needed (buf may be null to size only)
size_t msm_carve(const CurveVTable* vt, const MsmPlan& pl, size_t n, bool need_points, uint8_t* base, MsmBuffers* b,
                 uint8_t** scalars_dev, uint8_t** pts_in_dev, size_t pts_in_bytes, uint8_t** out_dev, size_t gmul = 1) {
    // gmul = 1: G1 (coordinates in Fp), 2: G2 (coordinates in Fp2: twice the bytes per point)
    const size_t aff_size = vt->aff_size * gmul, xyzz_size = vt->xyzz_size * gmul;
    const size_t n_sc = n;                 // scalars
    if (pl.glv) n *= 2;                    // points / digits / sorted entries: P_i and phi(P_i)
    size_t off = 0;
    auto take = [&](size_t bytes) { uint8_t* p = base ? base + off : nullptr; off += align_up(bytes); return p; };
    size_t nb = (size_t)pl.W * pl.B;
    b->points = need_points ? take(n * aff_size) : nullptr;
    b->digits = (uint32_t*)take((size_t)pl.W * n * 4);
    b->sorted = (uint32_t*)take((size_t)pl.W * n * 4);
    b->counts = (uint32_t*)take(nb * 4);
    b->offsets = (uint32_t*)take(nb * 4);
    b->cursor = (uint32_t*)take(nb * 4);
    b->perm = (uint32_t*)take(nb * 4);
    b->size_hist = (uint32_t*)take(2 * 1024 * 4);
    b->buckets = take(nb * xyzz_size);
    b->chunks = take((size_t)pl.W * pl.nchunks * xyzz_size);
    b->windows = take((size_t)pl.W * 9 * xyzz_size);      // W window sums + 8 partial sums per window
    b->max_heavy = (uint32_t)(((size_t)pl.W * n) / (size_t)pl.seg);     // extra segments of runs longer than pl.seg points
    b->heavy_n = (uint32_t*)take(4);
    b->heavy_items = take((size_t)b->max_heavy * 8);
    b->heavy_partial = take((size_t)b->max_heavy * xyzz_size);
    if (scalars_dev) *scalars_dev = take(n_sc * 32);
    if (pts_in_dev) *pts_in_dev = take(pts_in_bytes);
    if (out_dev) *out_dev = take(2 * (size_t)vt->fp_bytes * gmul);
    return off;
}

// MSM of host points/scalars [lo,hi) on one device -> one affine point written to out_host
int msm_host_range(const CurveInfo& ci, int dev, size_t lo, size_t hi, const void* pts, const void* resident_pts,
                   const void* scalars, void* out_host, uint32_t flags, const Bases* bs = nullptr, bool g2 = false) {
    const CurveVTable* vt = ci.vt;
    size_t m = hi - lo;
    const size_t gmul = g2 ? 2 : 1;
    size_t g1sz = 2 * (size_t)vt->fp_bytes * gmul;           // bytes of one point on the wire (G1 or G2)
    WsGuard g(dev);
    if (!g.w) return fail(B200_ERR_CUDA, "cannot create workspace on device %d", dev);
    Workspace& w = *g.w;
    CU(cudaSetDevice(dev));
    bool need_points = resident_pts == nullptr;
    // one-shot G1 MSM on a BLS12 curve: GLV split (the points are converted anyway, phi(P) is one more product each)
    MsmPlan pl = (need_points && !g2 && vt->glv && m >= B200_MSM_GLV_MIN) ? msm_plan_glv(m, vt->glv) : resident_plan(vt, bs, m);
    MsmBuffers b;
    uint8_t *d_sc = nullptr, *d_pin = nullptr, *d_out = nullptr;
    size_t need = msm_carve(vt, pl, m, need_points, nullptr, &b, &d_sc, need_points ? &d_pin : nullptr, m * g1sz, &d_out, gmul);
    if (int rc = w.reserve(need)) return rc;
    msm_carve(vt, pl, m, need_points, w.buf, &b, &d_sc, need_points ? &d_pin : nullptr, m * g1sz, &d_out, gmul);
    CU(cudaMemsetAsync(w.d_err, 0, sizeof(int), w.stream));
    const void* prepared = resident_pts;
    b.points_ready = nullptr;
    if (m) {
        if (need_points) {
            // the points (2/3 of the bytes) upload and convert on the copy stream while the main stream uploads the scalars
            // and runs the digit / scan / sort kernels, which need only the scalars; the bucket kernel waits on the event
            CU(cudaEventRecord(w.ev_start, w.stream));                     // orders the slab's reuse after earlier work
            CU(cudaStreamWaitEvent(w.copy_stream, w.ev_start, 0));
            CU(cudaMemcpyAsync(d_pin, (const uint8_t*)pts + lo * g1sz, m * g1sz, cudaMemcpyHostToDevice, w.copy_stream));
            if (g2) CU(vt->msm_points_g2(m, d_pin, b.points, kernel_flags(flags), w.d_err, w.copy_stream));
            else CU(vt->msm_points(m, d_pin, b.points, kernel_flags(flags), w.d_err, w.copy_stream, pl.glv));
            CU(cudaEventRecord(w.ev_points, w.copy_stream));
            b.points_ready = w.ev_points;
            prepared = b.points;
        }
        CU(cudaMemcpyAsync(d_sc, (const uint8_t*)scalars + lo * 32, m * 32, cudaMemcpyHostToDevice, w.stream));
        if (need_points) {
        } else {
            prepared = (const uint8_t*)resident_pts + lo * vt->aff_size;
        }
    }
    CU((g2 ? vt->msm_g2 : vt->msm)(m, prepared, d_sc, d_out, kernel_flags(flags), pl, b, w.stream));
    int h_err = 0;
    CU(cudaMemcpyAsync(out_host, d_out, g1sz, cudaMemcpyDeviceToHost, w.stream));
    CU(cudaMemcpyAsync(&h_err, w.d_err, sizeof(int), cudaMemcpyDeviceToHost, w.stream));
    CU(cudaStreamSynchronize(w.stream));
    if (h_err) return fail(B200_ERR_ENCODING, "input is not a canonical element encoding");
    return 0;
}

}  // namespace

// =================================================================================================
extern "C" {

int b200_init(uint32_t device_mask) {
    std::lock_guard<std::mutex> lk(g_mu);
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt == 0) {
        g_inited.store(false);
        return fail(B200_ERR_NOGPU, "no CUDA device available (%s); libb200math has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    g_devices.clear();
    for (int i = 0; i < cnt && i < 32; i++)
        if (device_mask == 0 || (device_mask >> i) & 1) g_devices.push_back(i);
    if (g_devices.empty()) return fail(B200_ERR_ARG, "device mask 0x%x selects none of the %d devices", device_mask, cnt);
    g_inited.store(true, std::memory_order_release);
    return 0;
}

void b200_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& kv : g_free_ws) {
        cudaSetDevice(kv.first);
        for (Workspace* w : kv.second) {
            if (w->buf) cudaFree(w->buf);
            if (w->d_err) cudaFree(w->d_err);
            if (w->stream) cudaStreamDestroy(w->stream);
            delete w;
        }
    }
    g_free_ws.clear();
    for (auto& kv : g_bases) { cudaSetDevice(kv.second.dev); cudaFree(kv.second.pts); }
    g_bases.clear();
    for (auto& kv : g_lines) { cudaSetDevice(kv.second.dev); cudaFree(kv.second.lines); cudaFree(kv.second.qinf); }
    g_lines.clear();
    g_inited.store(false);
}

const char* b200_last_error(void) { return t_err.c_str(); }

int b200_device_count(void) {
    if (ensure_init()) return 0;
    return (int)devices().size();
}

int b200_set_device(int device) {
    if (int rc = ensure_init()) return rc;
    int cnt = 0;
    cudaGetDeviceCount(&cnt);
    if (device < 0 || device >= cnt) return fail(B200_ERR_ARG, "device %d out of range (%d devices)", device, cnt);
    t_device = device;
    return 0;
}

int b200_set_stream(void* cuda_stream) {
    t_stream = (cudaStream_t)cuda_stream;
    return 0;
}

int b200_take_error(int* had_error) {
    if (int rc = ensure_init()) return rc;
    if (!had_error) return fail(B200_ERR_ARG, "null buffer");
    int dev = current_device();
    CU(cudaSetDevice(dev));
    int* flag = device_err_flag(dev);
    int h = 0;
    CU(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, t_stream));
    CU(cudaMemsetAsync(flag, 0, sizeof(int), t_stream));
    CU(cudaStreamSynchronize(t_stream));
    *had_error = h ? 1 : 0;
    if (h) t_err = "a B200_DEVICE_PTRS call rejected an input (not a canonical element encoding, or a row index out of range)";
    return 0;
}

int b200_fp_bytes(int curve) {
    CurveInfo ci;
    if (!curve_info(curve, &ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    return ci.vt->fp_bytes;
}

uint64_t b200_launch_count(void) { return g_launch_count.load(); }

static int pairing_common(int curve, int np, size_t n, const void* g1a, const void* g2a, const void* g1b,
                          const void* g2b, void* out, uint32_t flags) {
    if (int rc = ensure_init()) return rc;
    CurveInfo ci;
    if (!curve_info(curve, &ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (n == 0) return 0;
    if (!g1a || !g2a || !out || (np == 2 && (!g1b || !g2b))) return fail(B200_ERR_ARG, "null buffer");
    const CurveVTable* vt = ci.vt;
    uint32_t kf = kernel_flags(flags);
    if (ci.kilic) kf |= B200_FEXP;
    size_t fb = vt->fp_bytes, g1sz = 2 * fb, g2sz = 4 * fb;
    size_t osz = (flags & B200_OUT_UNITY_ONLY) ? 1 : 12 * fb;
    if (flags & B200_DEVICE_PTRS) {
        int dev = current_device();
        CU(cudaSetDevice(dev));
        CU(vt->pairing(np, n, (const uint8_t*)g1a, (const uint8_t*)g2a, (const uint8_t*)g1b, (const uint8_t*)g2b,
                       (uint8_t*)out, kf, device_err_flag(dev), t_stream));
        return 0;
    }
    auto devs = split_devices(n, 256);
    return run_split(devs, n, [&](int dev, size_t lo, size_t hi) {
        std::vector<Piece> ins = {{g1a, g1sz, nullptr}, {g2a, g2sz, nullptr}};
        if (np == 2) { ins.push_back({g1b, g1sz, nullptr}); ins.push_back({g2b, g2sz, nullptr}); }
        return staged_call(dev, lo, hi, ins, out, osz,
                           [&](size_t m, std::vector<Piece>& p, uint8_t* d_out, int* d_err, cudaStream_t s) {
                               return vt->pairing(np, m, p[0].dev, p[1].dev, np == 2 ? p[2].dev : nullptr,
                                                  np == 2 ? p[3].dev : nullptr, d_out, kf, d_err, s);
                           });
    });
}

int b200_pairing_batch(int curve, size_t n, const void* g1, const void* g2, void* gt_out, uint32_t flags) {
    return pairing_common(curve, 1, n, g1, g2, nullptr, nullptr, gt_out, flags);
}

int b200_pairing2_batch(int curve, size_t n, const void* g1a, const void* g2a, const void* g1b, const void* g2b,
                        void* gt_out, uint32_t flags) {
    return pairing_common(curve, 2, n, g1a, g2a, g1b, g2b, gt_out, flags);
}

int b200_fexp_batch(int curve, size_t n, const void* gt_in, void* gt_out, uint32_t flags) {
    if (int rc = ensure_init()) return rc;
    CurveInfo ci;
    if (!curve_info(curve, &ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (n == 0) return 0;
    if (!gt_in || !gt_out) return fail(B200_ERR_ARG, "null buffer");
    const CurveVTable* vt = ci.vt;
    // driver.Curve.FExp always exponentiates for gurvy ids and is the identity for kilic ids
    uint32_t kf = (kernel_flags(flags) & ~B200_FEXP) | (ci.kilic ? 0u : B200_FEXP);
    size_t gtsz = 12 * (size_t)vt->fp_bytes;
    size_t osz = (flags & B200_OUT_UNITY_ONLY) ? 1 : gtsz;
    if (flags & B200_DEVICE_PTRS) {
        int dev = current_device();
        CU(cudaSetDevice(dev));
        CU(vt->fexp(n, (const uint8_t*)gt_in, (uint8_t*)gt_out, kf, device_err_flag(dev), t_stream));
        return 0;
    }
    auto devs = split_devices(n, 256);
    return run_split(devs, n, [&](int dev, size_t lo, size_t hi) {
        std::vector<Piece> ins = {{gt_in, gtsz, nullptr}};
        return staged_call(dev, lo, hi, ins, gt_out, osz,
                           [&](size_t m, std::vector<Piece>& p, uint8_t* d_out, int* d_err, cudaStream_t s) {
                               return vt->fexp(m, p[0].dev, d_out, kf, d_err, s);
                           });
    });
}

int b200_g1_mul_batch(int curve, size_t n, const void* pts, const void* scalars, void* out, uint32_t flags) {
    if (int rc = ensure_init()) return rc;
    CurveInfo ci;
    if (!curve_info(curve, &ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (n == 0) return 0;
    if (!pts || !scalars || !out) return fail(B200_ERR_ARG, "null buffer");
    const CurveVTable* vt = ci.vt;
    uint32_t kf = kernel_flags(flags);
    size_t g1sz = 2 * (size_t)vt->fp_bytes;
    if (flags & B200_DEVICE_PTRS) {
        int dev = current_device();
        CU(cudaSetDevice(dev));
        CU(vt->g1_mul(n, (const uint8_t*)pts, (const uint8_t*)scalars, (uint8_t*)out, kf, device_err_flag(dev), t_stream));
        return 0;
    }
    auto devs = split_devices(n, 1024);
    return run_split(devs, n, [&](int dev, size_t lo, size_t hi) {
        std::vector<Piece> ins = {{pts, g1sz, nullptr}, {scalars, 32, nullptr}};
        return staged_call(dev, lo, hi, ins, out, g1sz,
                           [&](size_t m, std::vector<Piece>& p, uint8_t* d_out, int* d_err, cudaStream_t s) {
                               return vt->g1_mul(m, p[0].dev, p[1].dev, d_out, kf, d_err, s);
                           });
    });
}

int b200_g1_mul2_batch(int curve, size_t n, const void* P, const void* e, const void* Q, const void* f, void* out,
                       uint32_t flags) {
    if (int rc = ensure_init()) return rc;
    CurveInfo ci;
    if (!curve_info(curve, &ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (n == 0) return 0;
    if (!P || !e || !Q || !f || !out) return fail(B200_ERR_ARG, "null buffer");
    const CurveVTable* vt = ci.vt;
    uint32_t kf = kernel_flags(flags);
    size_t g1sz = 2 * (size_t)vt->fp_bytes;
    if (flags & B200_DEVICE_PTRS) {
        int dev = current_device();
        CU(cudaSetDevice(dev));
        CU(vt->g1_mul2(n, (const uint8_t*)P, (const uint8_t*)e, (const uint8_t*)Q, (const uint8_t*)f, (uint8_t*)out, kf,
                       device_err_flag(dev), t_stream));
        return 0;
    }
    auto devs = split_devices(n, 1024);
    return run_split(devs, n, [&](int dev, size_t lo, size_t hi) {
        std::vector<Piece> ins = {{P, g1sz, nullptr}, {e, 32, nullptr}, {Q, g1sz, nullptr}, {f, 32, nullptr}};
        return staged_call(dev, lo, hi, ins, out, g1sz,
                           [&](size_t m, std::vector<Piece>& p, uint8_t* d_out, int* d_err, cudaStream_t s) {
                               return vt->g1_mul2(m, p[0].dev, p[1].dev, p[2].dev, p[3].dev, d_out, kf, d_err, s);
                           });
    });
}

int b200_g1_sum(int curve, size_t n, const void* pts, void* out, uint32_t flags) {
    if (int rc = ensure_init()) return rc;
    CurveInfo ci;
    if (!curve_info(curve, &ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (!out || (n && !pts)) return fail(B200_ERR_ARG, "null buffer");
    const CurveVTable* vt = ci.vt;
    uint32_t kf = kernel_flags(flags);
    size_t g1sz = 2 * (size_t)vt->fp_bytes;
    int dev = current_device();
    CU(cudaSetDevice(dev));
    if (flags & B200_DEVICE_PTRS) {
        CU(vt->g1_sum(n, (const uint8_t*)pts, (uint8_t*)out, kf, device_err_flag(dev), t_stream));
        return 0;
    }
    WsGuard g(dev);
    if (!g.w) return fail(B200_ERR_CUDA, "cannot create workspace on device %d", dev);
    Workspace& w = *g.w;
    if (int rc = w.reserve(align_up(n * g1sz) + g1sz)) return rc;
    uint8_t* d_out = w.buf + align_up(n * g1sz);
    if (n) CU(cudaMemcpyAsync(w.buf, pts, n * g1sz, cudaMemcpyHostToDevice, w.stream));
    CU(cudaMemsetAsync(w.d_err, 0, sizeof(int), w.stream));
    CU(vt->g1_sum(n, w.buf, d_out, kf, w.d_err, w.stream));
    int h_err = 0;
    CU(cudaMemcpyAsync(out, d_out, g1sz, cudaMemcpyDeviceToHost, w.stream));
    CU(cudaMemcpyAsync(&h_err, w.d_err, sizeof(int), cudaMemcpyDeviceToHost, w.stream));
    CU(cudaStreamSynchronize(w.stream));
    if (h_err) return fail(B200_ERR_ENCODING, "input is not a canonical element encoding");
    return 0;
}

static int msm_device_ptrs(const CurveInfo& ci, size_t n, const void* pts, bool prepared, const void* scalars,
                           void* out, uint32_t flags, const Bases* bs = nullptr, bool g2 = false) {
    // device-pointer MSM: the call returns while its kernels are still queued, so the scratch slab cannot go back to
    // the pool at return.  It is kept per (calling thread, device, stream) -- two streams never share a slab, MSMs
    // issued on one stream are ordered by the stream -- and returned to the pool, after draining the stream, when the
    // thread ends (cgo threads come and go).
    struct Held {
        std::map<std::pair<int, cudaStream_t>, Workspace*> ws;
        ~Held() {
            for (auto& kv : ws) {
                if (!kv.second) continue;
                cudaSetDevice(kv.first.first);
                cudaStreamSynchronize(kv.first.second);
                ws_release(kv.second);
            }
        }
    };
    thread_local Held t_held;
    const CurveVTable* vt = ci.vt;
    int dev = current_device();
    CU(cudaSetDevice(dev));
    Workspace*& w = t_held.ws[std::make_pair(dev, t_stream)];
    if (!w) w = ws_acquire(dev);
    if (!w) return fail(B200_ERR_CUDA, "cannot create workspace on device %d", dev);
    bool need_points = !prepared;
    MsmPlan pl = (need_points && !g2 && vt->glv && n >= B200_MSM_GLV_MIN) ? msm_plan_glv(n, vt->glv) : resident_plan(vt, bs, n);
    MsmBuffers b;
    const size_t gmul = g2 ? 2 : 1;
    size_t need = msm_carve(vt, pl, n, need_points, nullptr, &b, nullptr, nullptr, 0, nullptr, gmul);
    if (need > w->cap) {
        CU(cudaStreamSynchronize(t_stream));     // earlier async work may still use the old slab
        if (int rc = w->reserve(need)) return rc;
    }
    msm_carve(vt, pl, n, need_points, w->buf, &b, nullptr, nullptr, 0, nullptr, gmul);
    const void* prep = pts;
    if (need_points && n) {
        if (g2) CU(vt->msm_points_g2(n, (const uint8_t*)pts, b.points, kernel_flags(flags), device_err_flag(dev), t_stream));
        else CU(vt->msm_points(n, (const uint8_t*)pts, b.points, kernel_flags(flags), device_err_flag(dev), t_stream, pl.glv));
        prep = b.points;
    }
    CU((g2 ? vt->msm_g2 : vt->msm)(n, prep, (const uint8_t*)scalars, (uint8_t*)out, kernel_flags(flags), pl, b, t_stream));
    return 0;
}

// G2 MSM: sum_i [k_i] Q_i over E'(Fp2).  One device per call: the caller shards ranges and combines the
// partial sums with b200_g2_sum, exactly as for G1.
int b200_g2_msm(int curve, size_t n, const void* pts, const void* scalars, void* out, uint32_t flags) {
    if (int rc = ensure_init()) return rc;
    CurveInfo ci;
    if (!curve_info(curve, &ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (!out || (n && (!pts || !scalars))) return fail(B200_ERR_ARG, "null buffer");
    if (flags & B200_DEVICE_PTRS) return msm_device_ptrs(ci, n, pts, false, scalars, out, flags, nullptr, true);
    return msm_host_range(ci, current_device(), 0, n, pts, nullptr, scalars, out, flags, nullptr, true);
}

int b200_g1_msm(int curve, size_t n, const void* pts, const void* scalars, void* out, uint32_t flags) {
    if (int rc = ensure_init()) return rc;
    CurveInfo ci;
    if (!curve_info(curve, &ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (!out || (n && (!pts || !scalars))) return fail(B200_ERR_ARG, "null buffer");
    const CurveVTable* vt = ci.vt;
    if (flags & B200_DEVICE_PTRS) return msm_device_ptrs(ci, n, pts, false, scalars, out, flags);
    auto devs = split_devices(n, (size_t)1 << 16);
    size_t g1sz = 2 * (size_t)vt->fp_bytes;
    if (devs.size() == 1) return msm_host_range(ci, devs[0], 0, n, pts, nullptr, scalars, out, flags);
    // range-split: one partial sum per GPU, then a tiny combine on the first GPU
    std::vector<uint8_t> partial(devs.size() * g1sz);
    size_t k = 0;
    std::vector<size_t> los, his;
    for (size_t d = 0; d < devs.size(); d++) { los.push_back(n * d / devs.size()); his.push_back(n * (d + 1) / devs.size()); }
    std::vector<std::thread> th;
    std::vector<int> rc(devs.size(), 0);
    std::vector<std::string> msg(devs.size());
    uint32_t pflags = (flags & ~B200_OUT_MONT) | B200_OUT_MONT;   // partials travel as MONT limbs
    for (k = 0; k < devs.size(); k++)
        th.emplace_back([&, k]() {
            rc[k] = msm_host_range(ci, devs[k], los[k], his[k], pts, nullptr, scalars, partial.data() + k * g1sz, pflags);
            if (rc[k]) msg[k] = t_err;
        });
    for (auto& t : th) t.join();
    for (k = 0; k < devs.size(); k++)
        if (rc[k]) { t_err = msg[k]; return rc[k]; }
    int saved = t_device;
    t_device = devs[0];
    int r = b200_g1_sum(curve, devs.size(), partial.data(), out, (flags & B200_OUT_MONT) | B200_IN_MONT);
    t_device = saved;
    return r;
}

}  // extern "C"

// element-wise batch over host or device buffers: ins[k] = (pointer, element bytes), one output element per item
template <class KFn>
static int elementwise_batch(int curve, size_t n, std::vector<Piece> ins, void* out, size_t out_fp_mult, uint32_t flags,
                             size_t min_per_dev, KFn kfn) {
    if (int rc = ensure_init()) return rc;
    CurveInfo ci;
    if (!curve_info(curve, &ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (n == 0) return 0;
    if (!out) return fail(B200_ERR_ARG, "null buffer");
    for (auto& p : ins)
        if (!p.host) return fail(B200_ERR_ARG, "null buffer");
    const CurveVTable* vt = ci.vt;
    const uint32_t kf = (kernel_flags(flags) & ~(B200_FEXP | B200_OUT_UNITY_ONLY)) | (flags & B200_NO_SUBGROUP_CHECK);
    for (auto& p : ins)
        if (p.elem != 32) p.elem *= (size_t)vt->fp_bytes;          // sizes are given in Fp units except 32-byte scalars
    const size_t osz = out_fp_mult ? out_fp_mult * (size_t)vt->fp_bytes : 1;      // 0: one verdict byte per item
    if (flags & B200_DEVICE_PTRS) {
        int dev = current_device();
        CU(cudaSetDevice(dev));
        for (auto& p : ins) p.dev = (uint8_t*)const_cast<void*>(p.host);
        CU(kfn(vt, n, ins, (uint8_t*)out, kf, device_err_flag(dev), t_stream));
        return 0;
    }
    auto devs = split_devices(n, min_per_dev);
    return run_split(devs, n, [&](int dev, size_t lo, size_t hi) {
        return staged_call(dev, lo, hi, ins, out, osz,
                           [&](size_t m, std::vector<Piece>& p, uint8_t* d_out, int* d_err, cudaStream_t s) {
                               return kfn(vt, m, p, d_out, kf, d_err, s);
                           });
    });
}

extern "C" {

int b200_g2_mul_batch(int curve, size_t n, const void* pts, const void* scalars, void* out, uint32_t flags) {
    return elementwise_batch(curve, n, {{pts, 4, nullptr}, {scalars, 32, nullptr}}, out, 4, flags, 1024,
                             [](const CurveVTable* vt, size_t m, std::vector<Piece>& p, uint8_t* d_out, uint32_t kf,
                                int* d_err, cudaStream_t s) { return vt->g2_mul(m, p[0].dev, p[1].dev, d_out, kf, d_err, s); });
}

int b200_g2_sum(int curve, size_t n, const void* pts, void* out, uint32_t flags) {
    if (int rc = ensure_init()) return rc;
    CurveInfo ci;
    if (!curve_info(curve, &ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (!out || (n && !pts)) return fail(B200_ERR_ARG, "null buffer");
    const CurveVTable* vt = ci.vt;
    uint32_t kf = kernel_flags(flags);
    size_t g2sz = 4 * (size_t)vt->fp_bytes;
    int dev = current_device();
    CU(cudaSetDevice(dev));
    if (flags & B200_DEVICE_PTRS) {
        CU(vt->g2_sum(n, (const uint8_t*)pts, (uint8_t*)out, kf, device_err_flag(dev), t_stream));
        return 0;
    }
    WsGuard g(dev);
    if (!g.w) return fail(B200_ERR_CUDA, "cannot create workspace on device %d", dev);
    Workspace& w = *g.w;
    if (int rc = w.reserve(align_up(n * g2sz) + g2sz)) return rc;
    if (n) CU(cudaMemcpyAsync(w.buf, pts, n * g2sz, cudaMemcpyHostToDevice, w.stream));
    uint8_t* d_out = w.buf + align_up(n * g2sz);
    CU(cudaMemsetAsync(w.d_err, 0, sizeof(int), w.stream));
    CU(vt->g2_sum(n, w.buf, d_out, kf, w.d_err, w.stream));
    int h_err = 0;
    CU(cudaMemcpyAsync(out, d_out, g2sz, cudaMemcpyDeviceToHost, w.stream));
    CU(cudaMemcpyAsync(&h_err, w.d_err, sizeof(int), cudaMemcpyDeviceToHost, w.stream));
    CU(cudaStreamSynchronize(w.stream));
    if (h_err) return fail(B200_ERR_ENCODING, "input is not a canonical element encoding");
    return 0;
}

int b200_g2_lines_upload(int curve, size_t n_q, const void* g2_pts, uint32_t flags, uint64_t* handle) {
    if (int rc = ensure_init()) return rc;
    CurveInfo ci;
    if (!curve_info(curve, &ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (!handle || !n_q || !g2_pts) return fail(B200_ERR_ARG, "null buffer");
    const CurveVTable* vt = ci.vt;
    int dev = current_device();
    CU(cudaSetDevice(dev));
    const size_t g2sz = 4 * (size_t)vt->fp_bytes, row = vt->lines_row_words() * sizeof(uint32_t);
    uint32_t* d_lines = nullptr;
    uint8_t* d_inf = nullptr;
    CU(cudaMalloc(&d_lines, n_q * row));
    if (cudaMalloc(&d_inf, n_q) != cudaSuccess) { cudaFree(d_lines); return fail(B200_ERR_CUDA, "cudaMalloc"); }
    int h_err = 0;
    auto cleanup = [&]() { cudaFree(d_lines); cudaFree(d_inf); };
    if (flags & B200_DEVICE_PTRS) {
        int* d_err = device_err_flag(dev);
        cudaMemsetAsync(d_err, 0, sizeof(int), t_stream);
        cudaError_t e = vt->lines_build(n_q, (const uint8_t*)g2_pts, d_lines, d_inf, kernel_flags(flags), d_err, t_stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, t_stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(t_stream);
        if (e != cudaSuccess) { cleanup(); return fail(B200_ERR_CUDA, "line table: %s", cudaGetErrorString(e)); }
    } else {
        WsGuard g(dev);
        if (!g.w) { cleanup(); return fail(B200_ERR_CUDA, "cannot create workspace on device %d", dev); }
        Workspace& w = *g.w;
        if (int rc = w.reserve(n_q * g2sz)) { cleanup(); return rc; }
        cudaError_t e = cudaMemcpyAsync(w.buf, g2_pts, n_q * g2sz, cudaMemcpyHostToDevice, w.stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(w.d_err, 0, sizeof(int), w.stream);
        if (e == cudaSuccess) e = vt->lines_build(n_q, w.buf, d_lines, d_inf, kernel_flags(flags), w.d_err, w.stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&h_err, w.d_err, sizeof(int), cudaMemcpyDeviceToHost, w.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(w.stream);
        if (e != cudaSuccess) { cleanup(); return fail(B200_ERR_CUDA, "line table: %s", cudaGetErrorString(e)); }
    }
    if (h_err) { cleanup(); return fail(B200_ERR_ENCODING, "input is not a canonical element encoding"); }
    std::lock_guard<std::mutex> lk(g_mu);
    uint64_t h = g_next_handle++;
    g_lines[h] = {curve, dev, n_q, d_lines, d_inf};
    *handle = h;
    return 0;
}

int b200_g2_lines_free(uint64_t handle) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_lines.find(handle);
    if (it == g_lines.end()) return fail(B200_ERR_ARG, "unknown line-table handle %llu", (unsigned long long)handle);
    cudaSetDevice(it->second.dev);
    cudaFree(it->second.lines);
    cudaFree(it->second.qinf);
    g_lines.erase(it);
    return 0;
}

static int pairing_fixed_common(uint64_t handle, int np, size_t n, const void* g1a, const uint32_t* qa, const void* g1b,
                                const uint32_t* qb, void* out, uint32_t flags) {
    if (int rc = ensure_init()) return rc;
    Lines ln;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_lines.find(handle);
        if (it == g_lines.end()) return fail(B200_ERR_ARG, "unknown line-table handle %llu", (unsigned long long)handle);
        ln = it->second;
    }
    if (n == 0) return 0;
    if (!g1a || !out || (np == 2 && !g1b)) return fail(B200_ERR_ARG, "null buffer");
    if (ln.n_q < (size_t)np && (!qa || (np == 2 && !qb))) return fail(B200_ERR_ARG, "default rows need %d table entries", np);
    CurveInfo ci;
    curve_info(ln.curve, &ci);
    const CurveVTable* vt = ci.vt;
    uint32_t kf = kernel_flags(flags);
    if (ci.kilic) kf |= B200_FEXP;
    const size_t g1sz = 2 * (size_t)vt->fp_bytes;
    const size_t osz = (flags & B200_OUT_UNITY_ONLY) ? 1 : 12 * (size_t)vt->fp_bytes;
    CU(cudaSetDevice(ln.dev));
    if (flags & B200_DEVICE_PTRS) {
        CU(vt->pairing_fixed(np, n, (const uint8_t*)g1a, qa, (const uint8_t*)g1b, qb, ln.lines, ln.qinf, (uint32_t)ln.n_q,
                             (uint8_t*)out, kf, device_err_flag(ln.dev), t_stream));
        return 0;
    }
    // row indices arrive from the host: reject out-of-range rows before they become device addresses
    for (size_t i = 0; i < n; i++)
        if ((qa && qa[i] >= ln.n_q) || (np == 2 && qb && qb[i] >= ln.n_q)) return fail(B200_ERR_ARG, "row index out of range");
    std::vector<Piece> ins = {{g1a, g1sz, nullptr}};
    if (np == 2) ins.push_back({g1b, g1sz, nullptr});
    const int ia = qa ? (int)ins.size() : -1;
    if (qa) ins.push_back({qa, 4, nullptr});
    const int ib = (np == 2 && qb) ? (int)ins.size() : -1;
    if (np == 2 && qb) ins.push_back({qb, 4, nullptr});
    return staged_call(ln.dev, 0, n, ins, out, osz,
                       [&](size_t m, std::vector<Piece>& p, uint8_t* d_out, int* d_err, cudaStream_t s) {
                           return vt->pairing_fixed(np, m, p[0].dev, ia >= 0 ? (const uint32_t*)p[ia].dev : nullptr,
                                                    np == 2 ? p[1].dev : nullptr,
                                                    ib >= 0 ? (const uint32_t*)p[ib].dev : nullptr, ln.lines, ln.qinf,
                                                    (uint32_t)ln.n_q, d_out, kf, d_err, s);
                       });
}

int b200_pairing_fixed_batch(uint64_t lines, size_t n, const void* g1, const uint32_t* q_idx, void* gt_out, uint32_t flags) {
    return pairing_fixed_common(lines, 1, n, g1, q_idx, nullptr, nullptr, gt_out, flags);
}

int b200_pairing2_fixed_batch(uint64_t lines, size_t n, const void* g1a, const uint32_t* qa_idx, const void* g1b,
                              const uint32_t* qb_idx, void* gt_out, uint32_t flags) {
    return pairing_fixed_common(lines, 2, n, g1a, qa_idx, g1b, qb_idx, gt_out, flags);
}

// op 0 decompress, 1 compress, 2 validate (launch.cuh: point_codec)
static int point_codec_batch(int curve, int g2, int op, size_t n, const void* in, void* out, uint32_t flags) {
    const size_t unc = g2 ? 4 : 2, cmp = g2 ? 2 : 1;
    if ((op == 0 && (flags & B200_IN_MONT)) || (op == 1 && (flags & B200_OUT_MONT)))
        return fail(B200_ERR_ARG, "compressed encodings have no MONT form");
    return elementwise_batch(curve, n, {{in, op == 0 ? cmp : unc, nullptr}}, out, op == 0 ? unc : (op == 1 ? cmp : 0), flags,
                             1024, [g2, op](const CurveVTable* vt, size_t m, std::vector<Piece>& p, uint8_t* d_out,
                                            uint32_t kf, int* d_err, cudaStream_t s) {
                                 return vt->point_codec(g2, op, m, p[0].dev, d_out, kf, d_err, s);
                             });
}
int b200_g1_decompress_batch(int curve, size_t n, const void* in, void* out, uint32_t flags) {
    return point_codec_batch(curve, 0, 0, n, in, out, flags);
}
int b200_g2_decompress_batch(int curve, size_t n, const void* in, void* out, uint32_t flags) {
    return point_codec_batch(curve, 1, 0, n, in, out, flags);
}
int b200_g1_compress_batch(int curve, size_t n, const void* in, void* out, uint32_t flags) {
    return point_codec_batch(curve, 0, 1, n, in, out, flags);
}
int b200_g2_compress_batch(int curve, size_t n, const void* in, void* out, uint32_t flags) {
    return point_codec_batch(curve, 1, 1, n, in, out, flags);
}
int b200_g1_validate_batch(int curve, size_t n, const void* in, void* ok_out, uint32_t flags) {
    return point_codec_batch(curve, 0, 2, n, in, ok_out, flags);
}
int b200_g2_validate_batch(int curve, size_t n, const void* in, void* ok_out, uint32_t flags) {
    return point_codec_batch(curve, 1, 2, n, in, ok_out, flags);
}

int b200_g1_normalize_batch(int curve, size_t n, const void* jac_mont, void* out, uint32_t flags) {
    if (flags & B200_IN_MONT) flags &= ~B200_IN_MONT;          // the input is always Montgomery limbs
    return elementwise_batch(curve, n, {{jac_mont, 3, nullptr}}, out, 2, flags, 4096,
                             [](const CurveVTable* vt, size_t m, std::vector<Piece>& p, uint8_t* d_out, uint32_t kf, int*,
                                cudaStream_t s) { return vt->g1_normalize(m, (const uint32_t*)p[0].dev, d_out, kf, s); });
}

int b200_hash_to_g1_batch(int curve, size_t n, const void* msgs, const uint64_t* offsets, const void* domain,
                          size_t domain_len, void* out, uint32_t flags) {
    if (int rc = ensure_init()) return rc;
    CurveInfo ci;
    if (!curve_info(curve, &ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    const CurveVTable* vt = ci.vt;
    if (!vt->hash_to_g1 || curve == B200_BLS12_377_GURVY)
        return fail(B200_ERR_ARG, "HashToG1 is built for the BLS12-381 curve ids (3, 5, 6, 7) only");
    if (domain_len > 255) return fail(B200_ERR_ARG, "invalid domain length");          // same refusal as custom.go:260-262
    if (n == 0) return 0;
    if (!offsets || !out || (domain_len && !domain)) return fail(B200_ERR_ARG, "null buffer");
    const int bbs = (curve == B200_BLS12_381_BBS || curve == B200_BLS12_381_BBS_GURVY) ? 1 : 0;
    const uint32_t kf = kernel_flags(flags) & B200_OUT_MONT;
    int dev = current_device();
    CU(cudaSetDevice(dev));
    if (flags & B200_DEVICE_PTRS) {
        CU(vt->hash_to_g1(bbs, n, (const uint8_t*)msgs, offsets, (const uint8_t*)domain, domain_len, (uint8_t*)out, kf, t_stream));
        return 0;
    }
    for (size_t i = 0; i < n; i++)
        if (offsets[i + 1] < offsets[i]) return fail(B200_ERR_ARG, "message offsets must be non-decreasing");
    const size_t total = (size_t)(offsets[n] - offsets[0]);
    if (total && !msgs) return fail(B200_ERR_ARG, "null buffer");
    WsGuard g(dev);
    if (!g.w) return fail(B200_ERR_CUDA, "cannot create workspace on device %d", dev);
    Workspace& w = *g.w;
    const size_t osz = n * 2 * (size_t)vt->fp_bytes;
    if (int rc = w.reserve(align_up(total + 1) + align_up((n + 1) * 8) + align_up(domain_len + 1) + align_up(osz))) return rc;
    uint8_t* d_msg = w.buf;
    uint64_t* d_off = (uint64_t*)(d_msg + align_up(total + 1));
    uint8_t* d_dst = (uint8_t*)d_off + align_up((n + 1) * 8);
    uint8_t* d_out = d_dst + align_up(domain_len + 1);
    // offsets are rebased to the staged copy, which starts at the first message
    std::vector<uint64_t> rel(n + 1);
    for (size_t i = 0; i <= n; i++) rel[i] = offsets[i] - offsets[0];
    if (total) CU(cudaMemcpyAsync(d_msg, (const uint8_t*)msgs + offsets[0], total, cudaMemcpyHostToDevice, w.stream));
    CU(cudaMemcpyAsync(d_off, rel.data(), (n + 1) * 8, cudaMemcpyHostToDevice, w.stream));
    if (domain_len) CU(cudaMemcpyAsync(d_dst, domain, domain_len, cudaMemcpyHostToDevice, w.stream));
    CU(vt->hash_to_g1(bbs, n, d_msg, d_off, d_dst, domain_len, d_out, kf, w.stream));
    CU(cudaMemcpyAsync(out, d_out, osz, cudaMemcpyDeviceToHost, w.stream));
    CU(cudaStreamSynchronize(w.stream));
    return 0;
}

int b200_gt_mul_batch(int curve, size_t n, const void* a, const void* b, void* out, uint32_t flags) {
    return elementwise_batch(curve, n, {{a, 12, nullptr}, {b, 12, nullptr}}, out, 12, flags, 256,
                             [](const CurveVTable* vt, size_t m, std::vector<Piece>& p, uint8_t* d_out, uint32_t kf,
                                int* d_err, cudaStream_t s) { return vt->gt_op(0, m, p[0].dev, p[1].dev, d_out, kf, d_err, s); });
}

int b200_gt_inv_batch(int curve, size_t n, const void* a, void* out, uint32_t flags) {
    return elementwise_batch(curve, n, {{a, 12, nullptr}}, out, 12, flags, 256,
                             [](const CurveVTable* vt, size_t m, std::vector<Piece>& p, uint8_t* d_out, uint32_t kf,
                                int* d_err, cudaStream_t s) { return vt->gt_op(1, m, p[0].dev, nullptr, d_out, kf, d_err, s); });
}

int b200_gt_exp_batch(int curve, size_t n, const void* a, const void* scalars, void* out, uint32_t flags) {
    return elementwise_batch(curve, n, {{a, 12, nullptr}, {scalars, 32, nullptr}}, out, 12, flags, 256,
                             [](const CurveVTable* vt, size_t m, std::vector<Piece>& p, uint8_t* d_out, uint32_t kf,
                                int* d_err, cudaStream_t s) { return vt->gt_op(2, m, p[0].dev, p[1].dev, d_out, kf, d_err, s); });
}

int b200_bases_upload(int curve, size_t n, const void* pts, uint32_t flags, uint64_t* handle) {
    if (int rc = ensure_init()) return rc;
    CurveInfo ci;
    if (!curve_info(curve, &ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (!handle || (n && !pts)) return fail(B200_ERR_ARG, "null buffer");
    const CurveVTable* vt = ci.vt;
    int dev = current_device();
    CU(cudaSetDevice(dev));
    struct DevMem {            // freed on every early return; release() hands the pointer to the handle table
        void* p = nullptr;
        ~DevMem() { if (p) cudaFree(p); }
        void* release() { void* q = p; p = nullptr; return q; }
    } mem;
    void*& d_pts = mem.p;
    MsmPlan tp = msm_plan(n ? n : 1, vt->scalar_bits);
    int rows = 1;
    if ((flags & B200_BASES_TABLES) && n && tp.W > 1) {
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        size_t need = (size_t)tp.W * n * vt->aff_size;
        if (need <= total_b / 4 && need <= free_b / 2) rows = tp.W;
    }
    CU(cudaMalloc(&d_pts, (n ? n : 1) * vt->aff_size * rows));
    size_t g1sz = 2 * (size_t)vt->fp_bytes;
    if (n) {
        if (flags & B200_DEVICE_PTRS) {
            CU(vt->msm_points(n, (const uint8_t*)pts, d_pts, kernel_flags(flags), device_err_flag(dev), t_stream, 0));
            if (rows > 1) CU(vt->msm_tables(n, tp, n, d_pts, t_stream));
            CU(cudaStreamSynchronize(t_stream));
        } else {
            WsGuard g(dev);
            if (!g.w) return fail(B200_ERR_CUDA, "cannot create workspace on device %d", dev);
            Workspace& w = *g.w;
            if (int rc = w.reserve(n * g1sz)) return rc;
            CU(cudaMemcpyAsync(w.buf, pts, n * g1sz, cudaMemcpyHostToDevice, w.stream));
            CU(cudaMemsetAsync(w.d_err, 0, sizeof(int), w.stream));
            CU(vt->msm_points(n, w.buf, d_pts, kernel_flags(flags), w.d_err, w.stream, 0));
            if (rows > 1) CU(vt->msm_tables(n, tp, n, d_pts, w.stream));
            int h_err = 0;
            CU(cudaMemcpyAsync(&h_err, w.d_err, sizeof(int), cudaMemcpyDeviceToHost, w.stream));
            CU(cudaStreamSynchronize(w.stream));
            if (h_err) return fail(B200_ERR_ENCODING, "input is not a canonical element encoding");
        }
    }
    std::lock_guard<std::mutex> lk(g_mu);
    uint64_t h = g_next_handle++;
    g_bases[h] = {curve, dev, n, mem.release(), rows > 1 ? tp.c : 0, rows > 1 ? tp.W : 0};
    *handle = h;
    return 0;
}

int b200_g1_msm_resident(uint64_t handle, size_t n, const void* scalars, void* out, uint32_t flags) {
    if (int rc = ensure_init()) return rc;
    Bases bs;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_bases.find(handle);
        if (it == g_bases.end()) return fail(B200_ERR_ARG, "unknown bases handle %llu", (unsigned long long)handle);
        bs = it->second;
    }
    if (n > bs.n) return fail(B200_ERR_ARG, "n=%zu exceeds the %zu resident bases", n, bs.n);
    if (!out || (n && !scalars)) return fail(B200_ERR_ARG, "null buffer");
    CurveInfo ci;
    curve_info(bs.curve, &ci);
    if (flags & B200_DEVICE_PTRS) {
        int saved = t_device;
        t_device = bs.dev;
        int r = msm_device_ptrs(ci, n, bs.pts, true, scalars, out, flags, &bs);
        t_device = saved;
        return r;
    }
    return msm_host_range(ci, bs.dev, 0, n, nullptr, bs.pts, scalars, out, flags, &bs);
}

int b200_bases_free(uint64_t handle) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_bases.find(handle);
    if (it == g_bases.end()) return fail(B200_ERR_ARG, "unknown bases handle %llu", (unsigned long long)handle);
    cudaSetDevice(it->second.dev);
    cudaFree(it->second.pts);
    g_bases.erase(it);
    return 0;
}

}  // extern "C"
