// Host side of the FDTD path: opaque handle, device memory, time loop, C ABI.
// Replaces the body of PModel.StaggeredFDTD_3D_with_relaxation
// (TranscranialModeling/BabelIntegrationBASE.py:2338-2365) below the Python boundary.
#include <stdarg.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include <map>
#include <mutex>
#include <thread>
#include <tuple>
#include <unistd.h>
#include "common.h"
#include "fdtd_kernels.cuh"
#include "fdtd_direct.cuh"
#include "fdtd_tma.cuh"
#include "fdtd_tma2.cuh"
#include "nccl_dyn.h"
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

static thread_local char g_err[1024] = "";
void bb_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char *bb_last_error(void) { return g_err; }
extern "C" const char *bb_version(void) { return "babelb200 0.1 (sm_100a)"; }

extern "C" int bb_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { bb_set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e)); return -1; }
    return n;
}
extern "C" int bb_device_name(int device, char *out, int out_len) {
    cudaDeviceProp prop;
    BB_CUDA(cudaGetDeviceProperties(&prop, device));
    snprintf(out, out_len, "%s", prop.name);
    return BB_OK;
}

extern "C" int bb_host_alloc(int64_t bytes, void **out) {
    BB_REQUIRE(out && bytes >= 0, "bad argument");
    BB_CUDA(cudaHostAlloc(out, (size_t)std::max<int64_t>(bytes, 16), cudaHostAllocDefault));
    return BB_OK;
}
extern "C" int bb_host_free(void *ptr) {
    if (ptr) BB_CUDA(cudaFreeHost(ptr));
    return BB_OK;
}

// ------------------------------------------------------------------------------------------
enum { CAT_STRESS = 0, CAT_PARTICLE, CAT_PML, CAT_OTHER, CAT_COUNT };

struct bb_fdtd {
    bb_fdtd_desc d;
    DevParams p;
    int label_bytes = 1;
    int nown = 0;
    cudaStream_t stream = nullptr, comm_stream = nullptr;
    bool own_stream = false;
    std::vector<void *> allocs;
    std::vector<size_t> alloc_bytes;
    int64_t device_bytes = 0;
    // sources
    int64_t nsrc_cells = 0, nsrc_boundary = 0;
    long long *src_cell = nullptr;
    int *src_row = nullptr;
    float *src_o[3] = {nullptr, nullptr, nullptr};
    float *srcfun = nullptr;  // [nt_src][nsrc]
    // SourceFunctions streamed during the run (bb_fdtd_set_source_functions_streamed): the caller's table, how much of it
    // (time samples) is on the device, and the copy stream / events that order each chunk before the step that reads it
    const void *sf_host = nullptr;
    int sf_f64 = 0, sf_chunk = 0, sf_uploaded = 0, sf_waited = 0;
    int64_t sf_stride = 0;
    cudaStream_t sf_stream = nullptr;
    cudaEvent_t sf_ev[2] = {nullptr, nullptr};
    int *bsrc_map = nullptr;  // [4][plane] boundary-plane source index map (multi-rank handles)
    // continuous-wave sources synthesised in the source kernel instead of read from srcfun (bb_fdtd_set_source_tones)
    float *tone_ac = nullptr, *tone_as = nullptr;     // [nsrc] A cos(phi), A sin(phi)
    std::vector<float> tone_es, tone_ec;              // [nt_src] ramp(n) sin(w t_n), ramp(n) cos(w t_n)
    // sensors
    int64_t nsensors = 0, nsamples = 0;
    long long *sensor_cell = nullptr;
    unsigned long long *sensor_findex = nullptr;   // 1-based Fortran-order indices (device-built table only)
    float *sensor_out = nullptr;  // [map][sample][sensor]
    int n_sensor_maps = 0, n_acc_maps = 0;
    int64_t step = 0;
    size_t xp_floats = 0, yp_floats = 0, zp_floats = 0;   // per part array
    int nxp = 0;                                          // planes of this slab inside the i-PML
    bool materials_set = false, maps_set = false, prepared = false;
    int *d_bad = nullptr;                                 // device flag: a label outside the material table was seen
    StressMaps smaps;
    ParticleMaps pmaps, pmaps16;                           // pmaps16: boxes of the 16-row particle tiles (fdtd_tma2.cuh)
    int chunk_override = 0, chunk_tail = 1, dbg_kernel = -1;
    // NVLink halo push
    bool peer_mode = false;
    unsigned long long *flags = nullptr;      // [2] flag words + [2] push counters (device)
    unsigned epoch = 1;
    void *ipc_opened[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // NCCL
    ncclComm_t comm = nullptr;
    cudaEvent_t ev_boundary = nullptr, ev_halo = nullptr;
    // timing
    cudaEvent_t ev_run0 = nullptr, ev_run1 = nullptr;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_cat;
    size_t ev_used = 0;
    bb_fdtd_stats stats;
};

// ------------------------------------------------------------------------------------------
// Device-memory cache.  A worker runs several simulations of the same grid in a row (forward, back-propagation, refocus:
// BabelIntegrationBASE.py:2338-2428), and cudaMalloc / cudaFree of the ~20 field arrays cost 0.2-0.5 s per simulation.
// Allocations of a destroyed handle therefore return to a per-device free list (exact-size reuse, at most
// BB_DEVICE_POOL_GB gigabytes, default 32) instead of to the driver; an allocation the driver cannot satisfy purges the
// list and retries.  bb_release_cached_memory() empties it.
// ------------------------------------------------------------------------------------------
struct DevPool { std::mutex mu; std::multimap<size_t, void *> free; size_t bytes = 0; };
static DevPool g_pool[64];
static size_t pool_cap() {
    static const size_t cap = [] { const char *e = getenv("BB_DEVICE_POOL_GB"); return (size_t)((e ? atof(e) : 32.0) * 1e9); }();
    return cap;
}
static void pool_purge(int dev) {
    DevPool &pl = g_pool[dev];
    std::lock_guard<std::mutex> lock(pl.mu);
    for (auto &kv : pl.free) cudaFree(kv.second);
    pl.free.clear();
    pl.bytes = 0;
}
static cudaError_t pool_alloc(int dev, void **ptr, size_t bytes) {
    DevPool &pl = g_pool[dev];
    {
        std::lock_guard<std::mutex> lock(pl.mu);
        auto it = pl.free.find(bytes);
        if (it != pl.free.end()) { *ptr = it->second; pl.free.erase(it); pl.bytes -= bytes; return cudaSuccess; }
    }
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); pool_purge(dev); e = cudaMalloc(ptr, bytes); }
    return e;
}
static void pool_free(int dev, void *ptr, size_t bytes) {
    if (!ptr) return;
    DevPool &pl = g_pool[dev];
    {
        std::lock_guard<std::mutex> lock(pl.mu);
        if (pl.bytes + bytes <= pool_cap()) { pl.free.emplace(bytes, ptr); pl.bytes += bytes; return; }
    }
    cudaFree(ptr);
}
extern "C" int bb_release_cached_memory(int device) {
    int ndev = bb_device_count();
    if (ndev < 0) return BB_ERR_CUDA;
    for (int d = 0; d < ndev && d < 64; d++) {
        if (device >= 0 && d != device) continue;
        BB_CUDA(cudaSetDevice(d));
        pool_purge(d);
    }
    return BB_OK;
}

static int dev_alloc(bb_fdtd *h, void **ptr, size_t bytes, bool zero = true) {
    if (bytes == 0) bytes = 16;
    BB_CUDA(pool_alloc(h->d.device, ptr, bytes));
    h->allocs.push_back(*ptr);
    h->alloc_bytes.push_back(bytes);
    h->device_bytes += (int64_t)bytes;
    if (zero) BB_CUDA(cudaMemsetAsync(*ptr, 0, bytes, h->stream));
    return BB_OK;
}

static int popcount32(uint32_t v) { return __builtin_popcount(v); }

// ------------------------------------------------------------------------------------------
// Staging: every bulk transfer between the caller's host arrays and the device goes through two fixed-size slots
// (device scratch + page-locked host buffer + event each), so the host-side copy of chunk c+1 overlaps the transfer and
// the conversion kernel of chunk c, and no entry point allocates or frees device memory per call.  One stager per
// device for the life of the process (a worker runs forward, back-propagation and refocus simulations in a row,
// BabelIntegrationBASE.py:2338-2428); the mutex serialises handles that share a device.
// ------------------------------------------------------------------------------------------
constexpr size_t BB_STAGE_BYTES = 32u << 20;
struct Stager {
    std::mutex mu;
    void *dev[2] = {nullptr, nullptr}, *host[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool ready = false;
};
static Stager g_stagers[64];

static int stager_get(int device, Stager **out) {
    BB_REQUIRE(device >= 0 && device < 64, "device ordinal %d", device);
    Stager &s = g_stagers[device];
    if (!s.ready) {
        for (int b = 0; b < 2; b++) {
            BB_CUDA(cudaMalloc(&s.dev[b], BB_STAGE_BYTES));
            BB_CUDA(cudaHostAlloc(&s.host[b], BB_STAGE_BYTES, cudaHostAllocDefault));
            BB_CUDA(cudaEventCreateWithFlags(&s.ev[b], cudaEventDisableTiming));
        }
        s.ready = true;
    }
    *out = &s;
    return BB_OK;
}

static bool host_is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// host -> host copy on a few threads (a single thread moves ~10 GB/s, PCIe 5 x16 takes 50)
static void parallel_memcpy(void *dst, const void *src, size_t n) {
    const size_t piece = 4u << 20;
    const int nt = (int)std::min<size_t>(4, n / piece);
    if (nt <= 1) { memcpy(dst, src, n); return; }
    std::vector<std::thread> th;
    const size_t per = (n / nt + 63) & ~(size_t)63;
    for (int t = 0; t < nt; t++) {
        const size_t o = (size_t)t * per, len = o >= n ? 0 : std::min(per, n - o);
        if (len) th.emplace_back([=] { memcpy((char *)dst + o, (const char *)src + o, len); });
    }
    for (auto &x : th) x.join();
}

// Upload `total` bytes from `src` in pieces of at most `piece` bytes (<= BB_STAGE_BYTES); consume(dev_ptr, offset, bytes)
// launches on h->stream whatever turns the staged piece into its final form.
template <class F>
static int staged_upload(bb_fdtd *h, const void *src, size_t total, size_t piece, F consume) {
    Stager *sg;
    int rc;
    if ((rc = stager_get(h->d.device, &sg))) return rc;
    std::lock_guard<std::mutex> lock(sg->mu);
    const bool pinned = host_is_pinned(src);
    int b = 0;
    for (size_t o = 0; o < total; o += piece, b ^= 1) {
        const size_t n = std::min(piece, total - o);
        BB_CUDA(cudaEventSynchronize(sg->ev[b]));          // the slot's previous transfer and kernel are done
        const void *from = (const char *)src + o;
        if (!pinned) { parallel_memcpy(sg->host[b], from, n); from = sg->host[b]; }
        BB_CUDA(cudaMemcpyAsync(sg->dev[b], from, n, cudaMemcpyHostToDevice, h->stream));
        if ((rc = consume(sg->dev[b], o, n))) return rc;
        BB_CUDA(cudaGetLastError());
        BB_CUDA(cudaEventRecord(sg->ev[b], h->stream));
    }
    BB_CUDA(cudaStreamSynchronize(h->stream));
    return BB_OK;
}

// Download `total` bytes to `dst` in pieces: produce(dev_ptr, offset, bytes) launches the kernel that writes the piece.
template <class F>
static int staged_download(bb_fdtd *h, void *dst, size_t total, size_t piece, F produce) {
    Stager *sg;
    int rc;
    if ((rc = stager_get(h->d.device, &sg))) return rc;
    std::lock_guard<std::mutex> lock(sg->mu);
    const bool pinned = host_is_pinned(dst);
    size_t pend_o[2] = {0, 0}, pend_n[2] = {0, 0};
    int b = 0;
    for (size_t o = 0; o < total; o += piece, b ^= 1) {
        const size_t n = std::min(piece, total - o);
        BB_CUDA(cudaEventSynchronize(sg->ev[b]));
        if (pend_n[b]) { parallel_memcpy((char *)dst + pend_o[b], sg->host[b], pend_n[b]); pend_n[b] = 0; }
        if ((rc = produce(sg->dev[b], o, n))) return rc;
        BB_CUDA(cudaGetLastError());
        if (pinned) BB_CUDA(cudaMemcpyAsync((char *)dst + o, sg->dev[b], n, cudaMemcpyDeviceToHost, h->stream));
        else { BB_CUDA(cudaMemcpyAsync(sg->host[b], sg->dev[b], n, cudaMemcpyDeviceToHost, h->stream)); pend_o[b] = o; pend_n[b] = n; }
        BB_CUDA(cudaEventRecord(sg->ev[b], h->stream));
    }
    BB_CUDA(cudaStreamSynchronize(h->stream));
    for (b = 0; b < 2; b++) if (pend_n[b]) parallel_memcpy((char *)dst + pend_o[b], sg->host[b], pend_n[b]);
    return BB_OK;
}

// scoped device allocation for the few temporaries that cannot be staged (freed on every return path)
struct DevTmp {
    void *p = nullptr;
    size_t n = 0;
    int dev = 0;
    explicit DevTmp(int device = 0) : dev(device) {}
    ~DevTmp() { release(); }
    void release() { if (p) pool_free(dev, p, n); p = nullptr; n = 0; }
    cudaError_t alloc(size_t bytes) { release(); n = bytes ? bytes : 16; return pool_alloc(dev, &p, n); }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// ------------------------------------------------------------------------------------------
// TMA descriptors (driver entry point resolved at run time: the library links only cudart)
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
    }
    return fn;
}

// 4-D map (k, j, plane, component) over `ncomp` consecutive row-major (d2, d1, d0) volumes of `esz`-byte elements:
// row pitch `rowpitch`, plane pitch `planepitch`, component pitch `comppitch` (elements); box (bw, bh, 1, depth);
// out-of-volume taps read zero
static int make_map4(CUtensorMap *m, const void *base, CUtensorMapDataType dt, int esz, long long d0, long long d1, long long d2, int ncomp,
                     long long rowpitch, long long planepitch, long long comppitch, int bw, int bh, int depth) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) { bb_set_error("cuTensorMapEncodeTiled is not available from this driver"); return BB_ERR_CUDA; }
    const cuuint64_t dims[4] = { (cuuint64_t)std::max<long long>(d0, 1), (cuuint64_t)std::max<long long>(d1, 1), (cuuint64_t)std::max<long long>(d2, 1), (cuuint64_t)ncomp };
    const cuuint64_t strides[3] = { (cuuint64_t)rowpitch * esz, (cuuint64_t)planepitch * esz, (cuuint64_t)comppitch * esz };
    const cuuint32_t box[4] = { (cuuint32_t)bw, (cuuint32_t)bh, 1, (cuuint32_t)depth };
    const cuuint32_t estr[4] = { 1, 1, 1, 1 };
    static const CUtensorMapL2promotion promo = [] {      // BB_L2PROMO=none|64|128|256 (experiments; default 128)
        const char *e = getenv("BB_L2PROMO");
        if (!e) return CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
        if (!strcmp(e, "none")) return CU_TENSOR_MAP_L2_PROMOTION_NONE;
        if (!strcmp(e, "64")) return CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
        if (!strcmp(e, "256")) return CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
        return CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    }();
    CUresult r = enc(m, dt, ncomp > 1 ? 4 : 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { bb_set_error("cuTensorMapEncodeTiled failed (%d) for box %dx%dx1x%d", (int)r, bw, bh, depth); return BB_ERR_CUDA; }
    return BB_OK;
}

static int make_tensor_maps(bb_fdtd *h) {
    using namespace tma;
    const DevParams &p = h->p;
    const CUtensorMapDataType F = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const long long vol = (long long)p.nloc * p.plane;
    int rc;
    // field groups over the pitched (nloc, n2, n3) volumes
    auto field = [&](CUtensorMap *m, const float *base, int ncomp, int bw, int bh, int depth) {
        return make_map4(m, base, F, 4, p.n3, p.n2, p.nloc, ncomp, p.pitch, p.plane, vol, bw, bh, depth);
    };
    if ((rc = field(&h->smaps.v3, p.V[0], 3, SW, SH, 3))) return rc;
    if ((rc = field(&h->pmaps.v3, p.V[0], 3, TX, TY, 3))) return rc;
    if ((rc = field(&h->smaps.s3, p.S[0], 6, TX, TY, 3))) return rc;
    if ((rc = field(&h->smaps.r3, p.R[0], 6, TX, TY, 3))) return rc;
    if ((rc = field(&h->smaps.pr, p.Pr, 1, TX, TY, 1))) return rc;
    if ((rc = field(&h->pmaps.sxx, p.S[0], 6, TX, TY, 1))) return rc;
    if ((rc = field(&h->pmaps.sh2, p.S[0], 6, SW, SH, 2))) return rc;
    if ((rc = field(&h->pmaps.sh3, p.S[0], 6, SW, SH, 3))) return rc;
    const bool u8 = h->label_bytes == 1;
    const int lw = u8 ? LabBox<uint8_t>::W : LabBox<uint16_t>::W;
    if ((rc = make_map4(&h->smaps.lab, p.lab, u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT16, h->label_bytes,
                        p.n3, p.n2, p.nloc + 1, 1, p.pitch, p.plane, 0, lw, LH, 1))) return rc;
    h->pmaps.lab = h->smaps.lab;
    // damped parts: components 0-2 normal stress parts, 3-4 shear stress parts, 5-7 velocity parts
    const long long ypl = (long long)p.nyrows * p.pitch, zpl = (long long)p.n2 * p.zpw;
    for (int depth = 2; depth <= 3; depth++) {
        CUtensorMap *mx = depth == 3 ? &h->smaps.xp3 : &h->smaps.xp2, *my = depth == 3 ? &h->smaps.yp3 : &h->smaps.yp2;
        CUtensorMap *mz = depth == 3 ? &h->smaps.zp3 : &h->smaps.zp2;
        if ((rc = make_map4(mx, p.XP[0], F, 4, p.n3, p.n2, h->nxp, BB_NPART, p.pitch, p.plane, (long long)h->xp_floats, TX, TY, depth))) return rc;
        if ((rc = make_map4(my, p.YP[0], F, 4, p.n3, p.nyrows, h->nown, BB_NPART, p.pitch, ypl, (long long)h->yp_floats, TX, TY, depth))) return rc;
        if ((rc = make_map4(mz, p.ZP[0], F, 4, p.zpw, p.n2, h->nown, BB_NPART, p.zpw, zpl, (long long)h->zp_floats, p.zbw, TY, depth))) return rc;
    }
    h->pmaps.xp3 = h->smaps.xp3; h->pmaps.yp3 = h->smaps.yp3; h->pmaps.zp3 = h->smaps.zp3;
    {   // the particle kernel on 16-row tiles: same tensors, taller boxes
        constexpr int R = 16, RH = R + 2 * HALO;
        ParticleMaps &m = h->pmaps16;
        if ((rc = field(&m.v3, p.V[0], 3, TX, R, 3))) return rc;
        if ((rc = field(&m.sxx, p.S[0], 6, TX, R, 1))) return rc;
        if ((rc = field(&m.sh2, p.S[0], 6, SW, RH, 2))) return rc;
        if ((rc = field(&m.sh3, p.S[0], 6, SW, RH, 3))) return rc;
        if ((rc = make_map4(&m.lab, p.lab, u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT16, h->label_bytes,
                            p.n3, p.n2, p.nloc + 1, 1, p.pitch, p.plane, 0, lw, R + 1, 1))) return rc;
        if ((rc = make_map4(&m.xp3, p.XP[0], F, 4, p.n3, p.n2, h->nxp, BB_NPART, p.pitch, p.plane, (long long)h->xp_floats, TX, R, 3))) return rc;
        if ((rc = make_map4(&m.yp3, p.YP[0], F, 4, p.n3, p.nyrows, h->nown, BB_NPART, p.pitch, ypl, (long long)h->yp_floats, TX, R, 3))) return rc;
        if ((rc = make_map4(&m.zp3, p.ZP[0], F, 4, p.zpw, p.n2, h->nown, BB_NPART, p.zpw, zpl, (long long)h->zp_floats, p.zbw, R, 3))) return rc;
    }
    if (p.acc_rms) { if ((rc = make_map4(&h->smaps.acc, p.acc_rms, F, 4, p.n3, p.n2, h->nown, 1, p.pitch, p.plane, 0, TX, TY, 1))) return rc; }
    else h->smaps.acc = h->smaps.pr;
    return BB_OK;
}

static int fdtd_create_body(bb_fdtd *h, const bb_fdtd_desc *d);
extern "C" void bb_fdtd_destroy(bb_fdtd *h);

extern "C" int bb_fdtd_create(const bb_fdtd_desc *d, bb_fdtd **out) {
    BB_REQUIRE(d && out, "null argument");
    BB_REQUIRE(d->n1 > 0 && d->n2 > 0 && d->n3 > 0, "bad grid %d %d %d", d->n1, d->n2, d->n3);
    BB_REQUIRE(d->pml >= 2 && 2 * d->pml < d->n1 && 2 * d->pml < d->n2 && 2 * d->pml < d->n3,
               "PML thickness %d must be >= 2 and leave an interior", d->pml);
    BB_REQUIRE(((long long)(d->i1 - d->i0) + 5) * d->n2 * (((long long)d->n3 + tma::TX - 1) / tma::TX * tma::TX) < (1ll << 32),
               "slab of %d planes x %d x %d exceeds 2^32 cells per field: split it over more GPUs", d->i1 - d->i0, d->n2, d->n3);
    BB_REQUIRE(d->pml <= tma::MAX_ZBW, "PML thickness %d above the supported maximum %d", d->pml, (int)tma::MAX_ZBW);
    BB_REQUIRE(d->i0 >= 0 && d->i1 <= d->n1 && d->i1 - d->i0 >= (d->nranks > 1 ? 4 : 1), "bad slab [%d,%d)", d->i0, d->i1);
    BB_REQUIRE(d->nmat >= 1 && d->nmat <= 32767, "nmat %d out of range", d->nmat);
    BB_REQUIRE(d->steps >= 0 && d->sensor_subsampling >= 1 && d->sensor_start >= 0, "bad time parameters");
    BB_REQUIRE(d->type_source >= 0 && d->type_source <= 3, "TypeSource %d not supported", d->type_source);
    BB_REQUIRE(d->sel_rms_peak >= 0 && d->sel_rms_peak <= 3, "SelRMSorPeak %d", d->sel_rms_peak);
    BB_REQUIRE((d->sel_maps_rms >> BB_MAP_COUNT) == 0 && (d->sel_maps_sensor >> BB_MAP_COUNT) == 0, "bad map mask");
    int ndev = bb_device_count();
    if (ndev <= 0) { if (ndev == 0) bb_set_error("no CUDA device (this library has no CPU fallback)"); return BB_ERR_CUDA; }
    BB_REQUIRE(d->device >= 0 && d->device < ndev, "device %d of %d", d->device, ndev);
    BB_CUDA(cudaSetDevice(d->device));
    bb_fdtd *h = new bb_fdtd();
    // everything created from here on is owned by the handle: a failure (most likely cudaMalloc on a grid that does not
    // fit) destroys it, so that the caller can retry with a smaller grid or more GPUs
    const int rc_create = fdtd_create_body(h, d);
    if (rc_create) { std::string msg = bb_last_error(); bb_fdtd_destroy(h); bb_set_error("%s", msg.c_str()); return rc_create; }
    *out = h;
    return BB_OK;
}

static int fdtd_create_body(bb_fdtd *h, const bb_fdtd_desc *d) {
    h->d = *d;
    memset(&h->stats, 0, sizeof(h->stats));
    BB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    {   // the halo exchange must win SMs against the interior kernel that is already queued: highest priority
        int lo = 0, hi = 0;
        BB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        BB_CUDA(cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, hi));
    }
    h->own_stream = true;
    BB_CUDA(cudaEventCreate(&h->ev_run0));
    BB_CUDA(cudaEventCreate(&h->ev_run1));
    BB_CUDA(cudaEventCreateWithFlags(&h->ev_boundary, cudaEventDisableTiming));
    BB_CUDA(cudaEventCreateWithFlags(&h->ev_halo, cudaEventDisableTiming));

    DevParams &p = h->p;
    memset(&p, 0, sizeof(p));
    p.n1 = d->n1; p.n2 = d->n2; p.n3 = d->n3; p.i0 = d->i0; p.i1 = d->i1; p.P = d->pml;
    p.pitch = (d->n3 + tma::TX - 1) / tma::TX * tma::TX;   // whole tiles: a row is a multiple of 256 bytes
    h->nown = d->i1 - d->i0;
    p.nloc = h->nown + 4;
    p.plane = (long long)p.n2 * p.pitch;
    p.dt = (float)d->dt;
    p.idt = (float)(1.0 / d->dt);
    p.mpml = d->mpml_ratio;
    p.nmat = d->nmat;
    h->label_bytes = d->nmat <= 127 ? 1 : 2;
    if (const char *e = getenv("BB_CHUNK")) h->chunk_override = atoi(e);
    if (const char *e = getenv("BB_CHUNK_TAIL")) h->chunk_tail = atoi(e);
    const size_t vol = (size_t)p.nloc * p.plane;
    int rc;
    // the components of a field group are contiguous (one 4-D TMA descriptor per group)
    if ((rc = dev_alloc(h, (void **)&p.V[0], 3 * vol * 4))) return rc;
    if ((rc = dev_alloc(h, (void **)&p.S[0], 6 * vol * 4))) return rc;
    if ((rc = dev_alloc(h, (void **)&p.R[0], 6 * vol * 4))) return rc;
    for (int c = 1; c < 3; c++) p.V[c] = p.V[0] + c * vol;
    for (int c = 1; c < 6; c++) { p.S[c] = p.S[0] + c * vol; p.R[c] = p.R[0] + c * vol; }
    if ((rc = dev_alloc(h, (void **)&p.Pr, vol * 4))) return rc;
    // label planes carry one extra zero plane so that (i+1) lookups of the last halo plane stay in bounds
    if ((rc = dev_alloc(h, (void **)&p.lab, (vol + p.plane) * h->label_bytes))) return rc;
    if ((rc = dev_alloc(h, (void **)&p.coef, sizeof(MatCoef) * d->nmat))) return rc;
    if ((rc = dev_alloc(h, (void **)&p.axI, sizeof(AxisCoef) * d->n1))) return rc;
    if ((rc = dev_alloc(h, (void **)&p.axJ, sizeof(AxisCoef) * d->n2))) return rc;
    if ((rc = dev_alloc(h, (void **)&p.axK, sizeof(AxisCoef) * d->n3))) return rc;
    p.ntk = p.pitch / tma::TX;
    p.ntj = (p.n2 + tma::TY - 1) / tma::TY;
    if ((rc = dev_alloc(h, (void **)&p.flags, (size_t)p.nloc * p.ntj * p.ntk))) return rc;
    // damped split parts of the PML shell of this slab
    const int P = d->pml;
    p.nxlo = std::max(0, std::min(P, d->i1) - d->i0);
    p.xhi_begin = std::min(std::max(d->n1 - P, d->i0), d->i1);
    // Y / Z parts are stored for whole tile rows / tile columns so that a TMA box maps 1:1 onto a tile
    p.nylo = (P + tma::TY - 1) / tma::TY; p.tjhi0 = (d->n2 - P) / tma::TY;
    if (p.tjhi0 < p.nylo) { p.nylo = p.ntj; p.tjhi0 = p.ntj; }   // tiny grids: every tile row is stored
    p.nyrows = (p.nylo + (p.ntj - p.tjhi0)) * tma::TY;
    p.zbw = (P + 3) / 4 * 4;
    p.zpw = 2 * p.zbw;
    h->nxp = p.nxlo + (d->i1 - p.xhi_begin);
    h->xp_floats = (size_t)h->nxp * p.plane;
    h->yp_floats = (size_t)h->nown * p.nyrows * p.pitch;
    h->zp_floats = (size_t)h->nown * p.n2 * p.zpw;
    h->xp_floats = std::max<size_t>(h->xp_floats, (size_t)p.plane);   // a slab without i-PML planes still needs a valid descriptor
    if ((rc = dev_alloc(h, (void **)&p.XP[0], BB_NPART * h->xp_floats * 4))) return rc;
    if ((rc = dev_alloc(h, (void **)&p.YP[0], BB_NPART * h->yp_floats * 4))) return rc;
    if ((rc = dev_alloc(h, (void **)&p.ZP[0], BB_NPART * h->zp_floats * 4))) return rc;
    for (int c = 1; c < BB_NPART; c++) { p.XP[c] = p.XP[0] + c * h->xp_floats; p.YP[c] = p.YP[0] + c * h->yp_floats; p.ZP[c] = p.ZP[0] + c * h->zp_floats; }
    // accumulators
    h->n_acc_maps = popcount32(d->sel_maps_rms);
    h->n_sensor_maps = popcount32(d->sel_maps_sensor);
    p.sel_maps = d->sel_maps_rms;
    p.sel_rms_peak = d->sel_rms_peak;
    p.acc_stride = (long long)h->nown * p.plane;
    if ((d->sel_rms_peak & 1) && h->n_acc_maps)
        if ((rc = dev_alloc(h, (void **)&p.acc_rms, (size_t)h->n_acc_maps * p.acc_stride * 4))) return rc;
    if ((d->sel_rms_peak & 2) && h->n_acc_maps)
        if ((rc = dev_alloc(h, (void **)&p.acc_peak, (size_t)h->n_acc_maps * p.acc_stride * 4))) return rc;
    for (int n = 0; n < d->steps; n++)
        if (n % d->sensor_subsampling == 0 && n / d->sensor_subsampling >= d->sensor_start) h->nsamples++;
    if (const char *e = getenv("BB_CTA_TIMING")) h->dbg_kernel = !strcmp(e, "stress") ? 0 : (!strcmp(e, "particle") ? 1 : -1);
    if (getenv("BB_CTA_TIMING")) { if ((rc = dev_alloc(h, (void **)&p.dbg, (size_t)4 * 8 * 65536))) return rc; }
    if ((rc = dev_alloc(h, (void **)&h->flags, 64))) return rc;
    p.flag_local = h->flags;
    p.push_count = reinterpret_cast<unsigned *>(h->flags + 2);
    if ((rc = dev_alloc(h, (void **)&h->d_bad, 16))) return rc;
    p.err = h->d_bad + 1;
    {   // BB_PEER_TIMEOUT_S: how long a boundary CTA waits for a neighbour's halo before the run is failed (default 60 s; 0 = forever)
        const char *e = getenv("BB_PEER_TIMEOUT_S");
        const double sec = e ? atof(e) : 60.0;
        p.peer_timeout_ns = sec > 0 ? (unsigned long long)(sec * 1e9) : 0ull;
        p.exp = getenv("BB_EXPERIMENT_HALO") ? atoi(getenv("BB_EXPERIMENT_HALO")) : 0;
    }
    if (d->kernel_variant != 1 && (rc = make_tensor_maps(h))) return rc;
    BB_CUDA(cudaStreamSynchronize(h->stream));
    return BB_OK;
}

extern "C" void bb_fdtd_destroy(bb_fdtd *h) {
    if (!h) return;
    cudaSetDevice(h->d.device);
    cudaDeviceSynchronize();
    // h->comm belongs to the process-wide cache (bb_fdtd_comm_init)
    for (void *m : h->ipc_opened) if (m) cudaIpcCloseMemHandle(m);
    for (size_t n = 0; n < h->allocs.size(); n++) pool_free(h->d.device, h->allocs[n], h->alloc_bytes[n]);
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    if (h->ev_run0) cudaEventDestroy(h->ev_run0);
    if (h->ev_run1) cudaEventDestroy(h->ev_run1);
    if (h->ev_boundary) cudaEventDestroy(h->ev_boundary);
    if (h->ev_halo) cudaEventDestroy(h->ev_halo);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
    if (h->sf_stream) cudaStreamDestroy(h->sf_stream);
    for (cudaEvent_t e : h->sf_ev) if (e) cudaEventDestroy(e);
    delete h;
}

extern "C" int bb_host_scatter_rows(void *out, const int64_t *rows, const void *data, int64_t nrows, int64_t row_bytes) {
    BB_REQUIRE(nrows == 0 || (out && rows && data), "null argument");
    BB_REQUIRE(nrows >= 0 && row_bytes > 0, "bad sizes");
    char *o = (char *)out;
    const char *d = (const char *)data;
    for (int64_t r = 0; r < nrows; r++) memcpy(o + rows[r] * row_bytes, d + r * row_bytes, (size_t)row_bytes);
    return BB_OK;
}

extern "C" int bb_host_scatter_runs(void *out, const void *data, const int64_t *dst_row, const int64_t *src_row, const int64_t *nrows,
                                    int64_t nruns, int64_t row_bytes) {
    BB_REQUIRE(nruns == 0 || (out && data && dst_row && src_row && nrows), "null argument");
    BB_REQUIRE(nruns >= 0 && row_bytes > 0, "bad sizes");
    char *o = (char *)out;
    const char *d = (const char *)data;
    for (int64_t r = 0; r < nruns; r++) {
        if (nrows[r] < 0 || dst_row[r] < 0 || src_row[r] < 0) { bb_set_error("negative run %lld", (long long)r); return BB_ERR_ARG; }
        memcpy(o + dst_row[r] * row_bytes, d + src_row[r] * row_bytes, (size_t)(nrows[r] * row_bytes));
    }
    return BB_OK;
}

// Nonzero entries of a uint32 volume (the caller's SourceMap: 46 656 of 18.4 M cells for CTX-500) on a few host threads:
// flat index and value of every nonzero entry, in order.  np.flatnonzero takes 40-60 ms for that volume, a third of the
// set-up phase of a simulation; this takes 3.  *count receives the number of nonzero entries; nothing is written when it
// exceeds `capacity` (call again with larger arrays).
extern "C" int bb_host_nonzero_u32(const uint32_t *a, int64_t n, int64_t *index, uint32_t *value, int64_t capacity, int64_t *count) {
    BB_REQUIRE(count && n >= 0 && (n == 0 || a) && capacity >= 0 && (capacity == 0 || (index && value)), "bad argument");
    const int nth = (int)std::max<int64_t>(1, std::min<int64_t>(8, n / (1 << 20)));
    const int64_t per = ((n + nth - 1) / nth + 7) & ~(int64_t)7;
    std::vector<std::vector<int64_t>> found(nth);
    auto body = [&](int t) {
        const int64_t b = std::min<int64_t>(n, t * per), e = std::min<int64_t>(n, b + per);
        std::vector<int64_t> &f = found[t];
        int64_t i = b;
        for (; i < e && (((uintptr_t)(a + i)) & 31); i++) if (a[i]) f.push_back(i);     // up to a 32-byte boundary
        for (; i + 8 <= e; i += 8) {                                                     // the volume is almost all zeros: 8 cells per test
            uint64_t w[4];
            memcpy(w, a + i, 32);
            if ((w[0] | w[1] | w[2] | w[3]) == 0) continue;
            for (int k = 0; k < 8; k++) if (a[i + k]) f.push_back(i + k);
        }
        for (; i < e; i++) if (a[i]) f.push_back(i);
    };
    if (nth == 1) body(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nth; t++) th.emplace_back(body, t);
        for (auto &x : th) x.join();
    }
    int64_t total = 0;
    for (const auto &f : found) total += (int64_t)f.size();
    *count = total;
    if (total > capacity) return BB_OK;
    int64_t o = 0;
    for (const auto &f : found) for (int64_t i : f) { index[o] = i; value[o] = a[i]; o++; }
    return BB_OK;
}

// LZ4 block decoder for babelbrain_b200/h5mini.py (the genuine H5pySimple writes its datasets as Blosc-LZ4 chunks; the
// pure-Python decoder manages a few MB/s).  Returns the number of bytes written, or -1 for a corrupt block.
extern "C" long long bb_host_lz4_decompress(const unsigned char *src, long long n, unsigned char *dst, long long cap) {
    if (!src || !dst || n < 0 || cap < 0) return -1;
    long long i = 0, o = 0;
    while (i < n) {
        const unsigned tok = src[i++];
        long long ll = tok >> 4;
        if (ll == 15) { unsigned b; do { if (i >= n) return -1; b = src[i++]; ll += b; } while (b == 255); }
        if (i + ll > n || o + ll > cap) return -1;
        memcpy(dst + o, src + i, (size_t)ll);
        i += ll; o += ll;
        if (i >= n) break;                          // the last sequence has no match
        if (i + 2 > n) return -1;
        const long long off = src[i] | (src[i + 1] << 8);
        i += 2;
        long long ml = tok & 15;
        if (ml == 15) { unsigned b; do { if (i >= n) return -1; b = src[i++]; ml += b; } while (b == 255); }
        ml += 4;
        if (off == 0 || off > o || o + ml > cap) return -1;
        for (long long k = 0; k < ml; k++) dst[o + k] = dst[o - off + k];     // may overlap: byte by byte
        o += ml;
    }
    return o;
}

extern "C" int bb_fdtd_set_stream(bb_fdtd *h, void *s) {
    BB_REQUIRE(h, "null handle");
    BB_CUDA(cudaSetDevice(h->d.device));
    BB_CUDA(cudaStreamSynchronize(h->stream));
    if (h->own_stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
    if (s) h->stream = (cudaStream_t)s;
    else { BB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true; }
    return BB_OK;
}

static void fill_axis(std::vector<AxisCoef> &t, int N, int P, const float *pml) {
    const float *D = pml, *DH = pml + (P + 1);     // damping at integer depth / half depth
    t.resize(N);
    for (int n = 0; n < N; n++) {
        AxisCoef &c = t[n];
        c.eI = c.eH = c.pad0 = c.pad1 = 0.0f;
        int d = 0, dh = -1;
        if (n < P) { d = P - n; dh = P - 1 - n; }
        else if (n >= N - P) { d = n - (N - P - 1); dh = d; }
        if (d > 0) c.eI = 0.5f * D[d];
        if (dh >= 0) c.eH = 0.5f * DH[dh];
        // backward difference landing on n, forward difference landing on n + 1/2
        if (n > 1 && n < N - 1) { c.cab = 1.125f; c.cbb = 1.0f / 24.0f; } else if (n > 0) { c.cab = 1.0f; c.cbb = 0.0f; } else { c.cab = c.cbb = 0.0f; }
        if (n > 0 && n < N - 2) { c.caf = 1.125f; c.cbf = 1.0f / 24.0f; } else if (n < N - 1) { c.caf = 1.0f; c.cbf = 0.0f; } else { c.caf = c.cbf = 0.0f; }
    }
}

extern "C" int bb_fdtd_set_materials(bb_fdtd *h, const float *table, const float *pml_table) {
    BB_REQUIRE(h && table && pml_table, "null argument");
    BB_CUDA(cudaSetDevice(h->d.device));
    // derived rows (double precision on the host, see MatCoef)
    std::vector<MatCoef> co(h->d.nmat);
    const double dt = h->d.dt;
    for (int m = 0; m < h->d.nmat; m++) {
        const float *r = table + (size_t)m * BB_NCOEF;
        const double M = r[0], G = r[1], L = r[2], B = r[3], tL = r[4], tS = r[5], ots = r[6], K = r[7];
        const double den = 1.0 + dt * 0.5 * ots, num = 1.0 - dt * 0.5 * ots;
        MatCoef &c = co[m];
        c.LM = (float)(M * (1.0 + tL)); c.Mi2 = (float)(2.0 * G * (1.0 + tS));
        c.LMCb = (float)(dt * M * tL * ots / den); c.MCb = (float)(dt * 2.0 * G * tS * ots / den);
        c.a = (float)(num / den); c.cs = (float)(dt * ots / den); c.K = (float)K;
        c.invG = G != 0.0 ? (float)(1.0 / G) : INFINITY;
        c.tauS = (float)tS; c.B = (float)B; c.M = (float)M; c.L = (float)L;
    }
    BB_CUDA(cudaMemcpyAsync((void *)h->p.coef, co.data(), sizeof(MatCoef) * h->d.nmat, cudaMemcpyHostToDevice, h->stream));
    std::vector<AxisCoef> ax[3];
    const int N[3] = { h->d.n1, h->d.n2, h->d.n3 };
    const AxisCoef *dst[3] = { h->p.axI, h->p.axJ, h->p.axK };
    for (int a = 0; a < 3; a++) {
        fill_axis(ax[a], N[a], h->d.pml, pml_table);
        BB_CUDA(cudaMemcpyAsync((void *)dst[a], ax[a].data(), sizeof(AxisCoef) * N[a], cudaMemcpyHostToDevice, h->stream));
    }
    BB_CUDA(cudaStreamSynchronize(h->stream));
    h->materials_set = true;
    h->prepared = false;
    return BB_OK;
}

extern "C" int bb_fdtd_set_maps(bb_fdtd *h, const uint32_t *material, const uint32_t *reflector) {
    BB_REQUIRE(h && material, "null argument");
    BB_CUDA(cudaSetDevice(h->d.device));
    const DevParams &p = h->p;
    const int glo = std::max(p.i0 - 2, 0), ghi = std::min(p.i1 + 2, p.n1);  // planes the host passes
    const long long nrows = (long long)(ghi - glo) * p.n2;
    const long long first_row = (long long)(glo - (p.i0 - 2)) * p.n2;  // local row where the host data starts
    BB_CUDA(cudaMemsetAsync(h->d_bad, 0, 4, h->stream));
    // whole rows per piece; with a reflector mask a piece carries the labels in its first half and the mask behind them
    const size_t row_bytes = (size_t)p.n3 * 4;
    const long long rows_per = std::max<long long>(1, (long long)(BB_STAGE_BYTES / (reflector ? 2 : 1) / row_bytes));
    BB_REQUIRE(row_bytes * (reflector ? 2 : 1) <= BB_STAGE_BYTES, "a row of %d labels does not fit the staging buffer", p.n3);
    const int bs = 256;
    int rc;
    if (!reflector) {
        rc = staged_upload(h, material, (size_t)nrows * row_bytes, (size_t)rows_per * row_bytes, [&](void *dev, size_t off, size_t n) {
            const long long r0 = (long long)(off / row_bytes), nr = (long long)(n / row_bytes);
            const unsigned grid = (unsigned)((nr * p.pitch + bs - 1) / bs);
            if (h->label_bytes == 1)
                label_convert_kernel<uint8_t><<<grid, bs, 0, h->stream>>>((const uint32_t *)dev, nullptr, (uint8_t *)p.lab + (first_row + r0) * p.pitch, nr, p.n3, p.pitch, h->d.nmat, h->d_bad);
            else
                label_convert_kernel<uint16_t><<<grid, bs, 0, h->stream>>>((const uint32_t *)dev, nullptr, (uint16_t *)p.lab + (first_row + r0) * p.pitch, nr, p.n3, p.pitch, h->d.nmat, h->d_bad);
            return BB_OK;
        });
        if (rc) return rc;
    } else {
        // the mask is rare (CT with air regions, BabelIntegrationBASE.py:2182-2190): upload it whole, then the labels in pieces
        DevTmp tmpr(h->d.device);
        BB_CUDA(tmpr.alloc((size_t)nrows * row_bytes));
        BB_CUDA(cudaMemcpyAsync(tmpr.p, reflector, (size_t)nrows * row_bytes, cudaMemcpyHostToDevice, h->stream));
        rc = staged_upload(h, material, (size_t)nrows * row_bytes, (size_t)rows_per * row_bytes, [&](void *dev, size_t off, size_t n) {
            const long long r0 = (long long)(off / row_bytes), nr = (long long)(n / row_bytes);
            const unsigned grid = (unsigned)((nr * p.pitch + bs - 1) / bs);
            const uint32_t *rf = tmpr.as<uint32_t>() + r0 * p.n3;
            if (h->label_bytes == 1)
                label_convert_kernel<uint8_t><<<grid, bs, 0, h->stream>>>((const uint32_t *)dev, rf, (uint8_t *)p.lab + (first_row + r0) * p.pitch, nr, p.n3, p.pitch, h->d.nmat, h->d_bad);
            else
                label_convert_kernel<uint16_t><<<grid, bs, 0, h->stream>>>((const uint32_t *)dev, rf, (uint16_t *)p.lab + (first_row + r0) * p.pitch, nr, p.n3, p.pitch, h->d.nmat, h->d_bad);
            return BB_OK;
        });
        if (rc) return rc;
    }
    int hbad = 0;
    BB_CUDA(cudaMemcpyAsync(&hbad, h->d_bad, 4, cudaMemcpyDeviceToHost, h->stream));
    BB_CUDA(cudaStreamSynchronize(h->stream));
    BB_REQUIRE(!hbad, "MaterialMap holds a label >= number of materials (%d)", h->d.nmat);
    h->maps_set = true;
    h->prepared = false;
    return BB_OK;
}

// global C-order cell index -> padded local index; returns -1 when the cell is not owned
static long long to_local(const DevParams &p, int64_t g) {
    const int64_t n23 = (int64_t)p.n2 * p.n3;
    const int i = (int)(g / n23);
    const int64_t r = g - (int64_t)i * n23;
    const int j = (int)(r / p.n3), k = (int)(r - (int64_t)j * p.n3);
    if (i < p.i0 || i >= p.i1) return -1;
    return ((long long)(i - p.i0 + 2) * p.n2 + j) * p.pitch + k;
}

extern "C" int bb_fdtd_set_source_cells(bb_fdtd *h, int64_t ncells, const int64_t *cell, const int32_t *row,
                                        const float *ox, const float *oy, const float *oz) {
    BB_REQUIRE(h && ncells >= 0, "bad argument");
    BB_REQUIRE(ncells == 0 || (cell && row && ox && oy && oz), "null source arrays");
    BB_CUDA(cudaSetDevice(h->d.device));
    const DevParams &p = h->p;
    // boundary-plane cells first (they are injected before the halo send), then interior cells
    std::vector<int64_t> order;
    order.reserve(ncells);
    std::vector<long long> loc(ncells);
    const bool multi = h->d.nranks > 1;
    int64_t nb = 0;
    for (int pass = 0; pass < 2; pass++)
        for (int64_t s = 0; s < ncells; s++) {
            if (pass == 0) {
                loc[s] = to_local(p, cell[s]);
                BB_REQUIRE(loc[s] >= 0, "source cell %lld is outside the slab [%d,%d)", (long long)cell[s], p.i0, p.i1);
                BB_REQUIRE(row[s] >= 0 && row[s] < h->d.nsrc, "source row %d out of range", row[s]);
            }
            const int ip = (int)(loc[s] / p.plane) - 2;
            // boundary = in one of the two planes next to a neighbour that exists (those planes are pushed to it)
            const bool boundary = multi && ((ip < 2 && h->d.rank > 0) || (ip >= h->nown - 2 && h->d.rank < h->d.nranks - 1));
            if ((pass == 0) == boundary) order.push_back(s);
            if (pass == 0 && boundary) nb++;
        }
    std::vector<long long> c2(ncells);
    std::vector<int> r2(ncells);
    std::vector<float> o2[3];
    for (int a = 0; a < 3; a++) o2[a].resize(ncells);
    for (int64_t t = 0; t < ncells; t++) {
        const int64_t s = order[t];
        c2[t] = loc[s]; r2[t] = row[s]; o2[0][t] = ox[s]; o2[1][t] = oy[s]; o2[2][t] = oz[s];
    }
    int rc;
    if ((rc = dev_alloc(h, (void **)&h->src_cell, ncells * 8, false))) return rc;
    if ((rc = dev_alloc(h, (void **)&h->src_row, ncells * 4, false))) return rc;
    for (int a = 0; a < 3; a++) if ((rc = dev_alloc(h, (void **)&h->src_o[a], ncells * 4, false))) return rc;
    if (ncells) {
        BB_CUDA(cudaMemcpy(h->src_cell, c2.data(), ncells * 8, cudaMemcpyHostToDevice));
        BB_CUDA(cudaMemcpy(h->src_row, r2.data(), ncells * 4, cudaMemcpyHostToDevice));
        for (int a = 0; a < 3; a++) BB_CUDA(cudaMemcpy(h->src_o[a], o2[a].data(), ncells * 4, cudaMemcpyHostToDevice));
    }
    h->nsrc_cells = ncells;
    h->nsrc_boundary = nb;
    if (multi) {
        // where the boundary-plane sources sit, for the half-step kernels that inject them themselves (DevParams::bsrc_map)
        std::vector<int> map((size_t)4 * p.plane, -1);
        for (int64_t t = 0; t < nb; t++) {
            const int ip = (int)(c2[t] / p.plane) - 2;
            const int b = ip < 2 ? ip : 2 + ip - (h->nown - 2);
            map[(size_t)b * p.plane + (size_t)(c2[t] % p.plane)] = (int)t;
        }
        if ((rc = dev_alloc(h, (void **)&h->bsrc_map, map.size() * 4, false))) return rc;
        BB_CUDA(cudaMemcpy(h->bsrc_map, map.data(), map.size() * 4, cudaMemcpyHostToDevice));
    }
    return BB_OK;
}

extern "C" int bb_fdtd_set_source_functions(bb_fdtd *h, const void *data, int is_f64, int64_t row_stride) {
    BB_REQUIRE(h && data, "null argument");
    BB_CUDA(cudaSetDevice(h->d.device));
    const int nsrc = h->d.nsrc, nt = h->d.nt_src;
    BB_REQUIRE(nsrc > 0 && nt > 0 && row_stride >= nt, "bad SourceFunctions shape");
    int rc;
    if (!h->srcfun) if ((rc = dev_alloc(h, (void **)&h->srcfun, (size_t)nsrc * nt * 4, false))) return rc;
    const size_t esz = is_f64 ? 8 : 4, row_bytes = (size_t)row_stride * esz;
    BB_REQUIRE(row_bytes <= BB_STAGE_BYTES, "a SourceFunctions row of %lld samples does not fit the staging buffer", (long long)row_stride);
    const dim3 blk(32, 8);
    // the caller's (nsrc, nt) float64 matrix (0.95 GB for CTX-500) streams through the two staging slots in groups of
    // whole rows; each group is converted and transposed into [nt][nsrc] while the next one is being copied
    const int64_t rows_per = std::max<int64_t>(1, (int64_t)(BB_STAGE_BYTES / row_bytes));
    const size_t total = ((size_t)(nsrc - 1) * row_stride + nt) * esz;
    return staged_upload(h, data, total, (size_t)rows_per * row_bytes, [&](void *dev, size_t off, size_t n) {
        const int64_t s0 = (int64_t)(off / row_bytes);
        const int64_t ns = std::min<int64_t>(rows_per, nsrc - s0);
        const dim3 grid((nt + 31) / 32, (unsigned)((ns + 31) / 32));
        if (is_f64) srcfun_transpose_kernel<double><<<grid, blk, 0, h->stream>>>((const double *)dev, row_stride, h->srcfun + s0, (int)ns, nt, nsrc);
        else srcfun_transpose_kernel<float><<<grid, blk, 0, h->stream>>>((const float *)dev, row_stride, h->srcfun + s0, (int)ns, nt, nsrc);
        return BB_OK;
    });
}

// ---- SourceFunctions streamed in time chunks while the time loop runs
// The dense table is 0.95 GB of float64 for CTX-500 and the time loop only ever needs the samples of the step it is at, so
// the upload does not have to finish before the first step: chunk c (sf_chunk consecutive samples of every row) is gathered
// into a page-locked staging slot by the host thread that launches the time loop (it runs hundreds of launches ahead of
// the GPU), copied and transposed on a second stream, and the time loop waits on the chunk's event at the first step
// that reads it.  One chunk is kept in flight ahead of the one in use.
static int upload_source_chunk(bb_fdtd *h) {
    const int nsrc = h->d.nsrc, nt = h->d.nt_src;
    const int t0 = h->sf_uploaded, T = std::min(h->sf_chunk, nt - t0);
    if (T <= 0) return BB_OK;
    const size_t esz = h->sf_f64 ? 8 : 4, seg = (size_t)T * esz;
    Stager *sg;
    int rc;
    if ((rc = stager_get(h->d.device, &sg))) return rc;
    std::lock_guard<std::mutex> lock(sg->mu);
    const int b = (t0 / h->sf_chunk) & 1;
    BB_CUDA(cudaEventSynchronize(sg->ev[b]));          // the slot's previous transfer and kernel are done
    {   // rows' segments [t0, t0 + T) -> dense (nsrc, T) block in the staging slot, on a few threads
        const char *src = (const char *)h->sf_host + (size_t)t0 * esz;
        char *dst = (char *)sg->host[b];
        const size_t rstride = (size_t)h->sf_stride * esz;
        const int nth = (size_t)nsrc * seg >= (8u << 20) ? 4 : 1;
        auto part = [=](int r0, int r1) { for (int r = r0; r < r1; r++) memcpy(dst + (size_t)r * seg, src + (size_t)r * rstride, seg); };
        if (nth == 1) part(0, nsrc);
        else {
            std::vector<std::thread> th;
            const int per = (nsrc + nth - 1) / nth;
            for (int t = 0; t < nth; t++) th.emplace_back(part, std::min(nsrc, t * per), std::min(nsrc, (t + 1) * per));
            for (auto &x : th) x.join();
        }
    }
    BB_CUDA(cudaMemcpyAsync(sg->dev[b], sg->host[b], (size_t)nsrc * seg, cudaMemcpyHostToDevice, h->sf_stream));
    const dim3 blk(32, 8), grid((T + 31) / 32, (unsigned)((nsrc + 31) / 32));
    float *out = h->srcfun + (size_t)t0 * nsrc;
    if (h->sf_f64) srcfun_transpose_kernel<double><<<grid, blk, 0, h->sf_stream>>>((const double *)sg->dev[b], T, out, nsrc, T, nsrc);
    else srcfun_transpose_kernel<float><<<grid, blk, 0, h->sf_stream>>>((const float *)sg->dev[b], T, out, nsrc, T, nsrc);
    BB_CUDA(cudaGetLastError());
    BB_CUDA(cudaEventRecord(sg->ev[b], h->sf_stream));
    BB_CUDA(cudaEventRecord(h->sf_ev[b], h->sf_stream));
    h->sf_uploaded = t0 + T;
    return BB_OK;
}

// before the kernels of time step n are launched: its chunk is on its way, the next one too, and the time loop's stream
// waits for the chunk of n if it has not yet
static int ensure_source_samples(bb_fdtd *h, int n) {
    if (!h->sf_host || n >= h->d.nt_src) return BB_OK;
    int rc;
    const int want = std::min(h->d.nt_src, (n / h->sf_chunk + 2) * h->sf_chunk);
    while (h->sf_uploaded < want) if ((rc = upload_source_chunk(h))) return rc;
    const int c = n / h->sf_chunk;
    if (c >= h->sf_waited) {
        BB_CUDA(cudaStreamWaitEvent(h->stream, h->sf_ev[c & 1], 0));
        h->sf_waited = c + 1;
    }
    return BB_OK;
}

extern "C" int bb_fdtd_set_source_functions_streamed(bb_fdtd *h, const void *data, int is_f64, int64_t row_stride) {
    BB_REQUIRE(h && data, "null argument");
    const int nsrc = h->d.nsrc, nt = h->d.nt_src;
    BB_REQUIRE(nsrc > 0 && nt > 0 && row_stride >= nt, "bad SourceFunctions shape");
    const size_t esz = is_f64 ? 8 : 4;
    const int64_t chunk = (int64_t)(BB_STAGE_BYTES / ((size_t)nsrc * esz));
    // many short rows (H317: 331 776 sources) would make chunks of a few samples: the whole-row upload serves those
    if (chunk < 64 || getenv("BB_SOURCES_UPFRONT")) return bb_fdtd_set_source_functions(h, data, is_f64, row_stride);
    BB_CUDA(cudaSetDevice(h->d.device));
    int rc;
    if (!h->srcfun) if ((rc = dev_alloc(h, (void **)&h->srcfun, (size_t)nsrc * nt * 4, false))) return rc;
    if (!h->sf_stream) {
        BB_CUDA(cudaStreamCreateWithFlags(&h->sf_stream, cudaStreamNonBlocking));
        for (cudaEvent_t &e : h->sf_ev) BB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    BB_CUDA(cudaStreamSynchronize(h->stream));          // srcfun may still be read by an earlier run of this handle
    h->sf_host = data; h->sf_f64 = is_f64; h->sf_stride = row_stride;
    h->sf_chunk = (int)std::min<int64_t>(chunk, 512);
    h->sf_uploaded = 0; h->sf_waited = 0;
    return BB_OK;
}

extern "C" int bb_fdtd_set_source_tones(bb_fdtd *h, const float *a_cos, const float *a_sin, const float *env_sin, const float *env_cos) {
    BB_REQUIRE(h && a_cos && a_sin && env_sin && env_cos, "null argument");
    BB_REQUIRE(!h->srcfun, "SourceFunctions were already set as a table");
    BB_CUDA(cudaSetDevice(h->d.device));
    const int nsrc = h->d.nsrc, nt = h->d.nt_src;
    BB_REQUIRE(nsrc > 0 && nt > 0, "bad SourceFunctions shape");
    int rc;
    if (!h->tone_ac) {
        if ((rc = dev_alloc(h, (void **)&h->tone_ac, (size_t)nsrc * 4, false))) return rc;
        if ((rc = dev_alloc(h, (void **)&h->tone_as, (size_t)nsrc * 4, false))) return rc;
    }
    BB_CUDA(cudaMemcpy(h->tone_ac, a_cos, (size_t)nsrc * 4, cudaMemcpyHostToDevice));
    BB_CUDA(cudaMemcpy(h->tone_as, a_sin, (size_t)nsrc * 4, cudaMemcpyHostToDevice));
    h->tone_es.assign(env_sin, env_sin + nt);
    h->tone_ec.assign(env_cos, env_cos + nt);
    return BB_OK;
}

extern "C" int bb_fdtd_set_sensors(bb_fdtd *h, int64_t nsensors, const int64_t *cell) {
    BB_REQUIRE(h && nsensors >= 0 && (nsensors == 0 || cell), "bad argument");
    BB_CUDA(cudaSetDevice(h->d.device));
    std::vector<long long> loc(nsensors);
    for (int64_t s = 0; s < nsensors; s++) {
        loc[s] = to_local(h->p, cell[s]);
        BB_REQUIRE(loc[s] >= 0, "sensor cell %lld is outside the slab", (long long)cell[s]);
    }
    int rc;
    if ((rc = dev_alloc(h, (void **)&h->sensor_cell, nsensors * 8, false))) return rc;
    if (nsensors) BB_CUDA(cudaMemcpy(h->sensor_cell, loc.data(), nsensors * 8, cudaMemcpyHostToDevice));
    if ((rc = dev_alloc(h, (void **)&h->sensor_out, (size_t)h->n_sensor_maps * h->nsamples * nsensors * 4))) return rc;
    BB_CUDA(cudaStreamSynchronize(h->stream));
    h->nsensors = nsensors;
    return BB_OK;
}

// ------------------------------------------------------------------------------------------
// sensor table built on the device (stream compaction in IndexSensorMap order)
// ------------------------------------------------------------------------------------------
struct SensorPred {
    const uint32_t *sm;   // (nown, n2, n3) C-order planes of the caller's SensorMap
    int nown, n2, n3;
    // t enumerates the slab in Fortran order (i fastest): the order of IndexSensorMap
    __device__ bool operator()(long long t) const {
        const int il = (int)(t % nown);
        const long long r = t / nown;
        const int j = (int)(r % n2), k = (int)(r / n2);
        return sm[((long long)il * n2 + j) * n3 + k] != 0;
    }
};

__global__ void sensor_finish_kernel(const long long *__restrict__ sel, long long n, DevParams p, long long *__restrict__ cell,
                                     unsigned long long *__restrict__ findex) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const long long t = sel[s];
    const int nown = p.i1 - p.i0;
    const int il = (int)(t % nown);
    const long long r = t / nown;
    const int j = (int)(r % p.n2), k = (int)(r / p.n2);
    cell[s] = ((long long)(il + 2) * p.n2 + j) * p.pitch + k;
    findex[s] = (unsigned long long)(il + p.i0) + (unsigned long long)j * p.n1 + (unsigned long long)k * p.n1 * p.n2 + 1ull;
}

__global__ void narrow_index_kernel(const unsigned long long *__restrict__ in, uint32_t *__restrict__ out, long long n) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) out[s] = (uint32_t)in[s];
}

extern "C" int bb_fdtd_set_sensor_map(bb_fdtd *h, const uint32_t *sensor_map, int64_t *nsensors) {
    BB_REQUIRE(h && sensor_map && nsensors, "null argument");
    BB_CUDA(cudaSetDevice(h->d.device));
    const DevParams &p = h->p;
    const long long total = (long long)h->nown * p.n2 * p.n3;
    // the selection needs the whole slab of the map on the device (Fortran-order enumeration of a C-order volume);
    // temporaries are scoped: every return path frees them
    DevTmp dmap(h->d.device), dsel(h->d.device), dcount(h->d.device), tmp(h->d.device);
    BB_CUDA(dmap.alloc((size_t)total * 4));
    {   // upload through the page-locked staging slots (the caller's map is pageable)
        int rcu = staged_upload(h, sensor_map, (size_t)total * 4, BB_STAGE_BYTES, [&](void *dev, size_t off, size_t n) {
            BB_CUDA(cudaMemcpyAsync((char *)dmap.p + off, dev, n, cudaMemcpyDeviceToDevice, h->stream));
            return BB_OK;
        });
        if (rcu) return rcu;
    }
    BB_CUDA(dsel.alloc((size_t)total * 8));   // worst case: every voxel is a sensor
    BB_CUDA(dcount.alloc(8));
    const SensorPred pred{dmap.as<uint32_t>(), h->nown, p.n2, p.n3};
    long long found = 0;
    const long long piece = 1ll << 30;               // keep each selection inside 32-bit item counts
    size_t tmp_bytes = 0;
    for (long long t0 = 0; t0 < total; t0 += piece) {
        const int n = (int)std::min(piece, total - t0);
        thrust::counting_iterator<long long> first(t0);
        size_t need = 0;
        BB_CUDA(cub::DeviceSelect::If(nullptr, need, first, dsel.as<long long>() + found, dcount.as<long long>(), n, pred, h->stream));
        if (need > tmp_bytes) { BB_CUDA(tmp.alloc(need)); tmp_bytes = need; }
        BB_CUDA(cub::DeviceSelect::If(tmp.p, need, first, dsel.as<long long>() + found, dcount.as<long long>(), n, pred, h->stream));
        long long c = 0;
        BB_CUDA(cudaMemcpyAsync(&c, dcount.p, 8, cudaMemcpyDeviceToHost, h->stream));
        BB_CUDA(cudaStreamSynchronize(h->stream));
        found += c;
    }
    int rc;
    if ((rc = dev_alloc(h, (void **)&h->sensor_cell, (size_t)found * 8, false))) return rc;
    if ((rc = dev_alloc(h, (void **)&h->sensor_findex, (size_t)found * 8, false))) return rc;
    if (found) {
        sensor_finish_kernel<<<(unsigned)((found + 255) / 256), 256, 0, h->stream>>>(dsel.as<long long>(), found, p, h->sensor_cell, h->sensor_findex);
        BB_CUDA(cudaGetLastError());
    }
    if ((rc = dev_alloc(h, (void **)&h->sensor_out, (size_t)h->n_sensor_maps * h->nsamples * found * 4))) return rc;
    BB_CUDA(cudaStreamSynchronize(h->stream));
    h->nsensors = found;
    *nsensors = found;
    return BB_OK;
}

extern "C" int bb_fdtd_get_sensor_index(bb_fdtd *h, void *out, int elem_bytes) {
    BB_REQUIRE(h && out && (elem_bytes == 4 || elem_bytes == 8), "bad argument");
    BB_REQUIRE(h->sensor_findex || h->nsensors == 0, "the sensor table was not built with bb_fdtd_set_sensor_map");
    BB_CUDA(cudaSetDevice(h->d.device));
    if (h->nsensors == 0) return BB_OK;
    if (elem_bytes == 4) BB_REQUIRE((long long)h->p.n1 * h->p.n2 * h->p.n3 < (1ll << 32), "grid too large for 32-bit sensor indices");
    const size_t per = BB_STAGE_BYTES / elem_bytes;
    return staged_download(h, out, (size_t)h->nsensors * elem_bytes, per * elem_bytes, [&](void *dev, size_t off, size_t n) {
        const long long s0 = (long long)(off / elem_bytes), ns = (long long)(n / elem_bytes);
        if (elem_bytes == 8) BB_CUDA(cudaMemcpyAsync(dev, h->sensor_findex + s0, n, cudaMemcpyDeviceToDevice, h->stream));
        else narrow_index_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, h->stream>>>(h->sensor_findex + s0, (uint32_t *)dev, ns);
        return BB_OK;
    });
}

extern "C" int bb_nccl_unique_id(char *out128) {
    BB_REQUIRE(out128, "null argument");
    if (!nccl_api().ok) { bb_set_error("NCCL not available: %s", nccl_api().err.c_str()); return BB_ERR_NCCL; }
    ncclUniqueId id;
    ncclResult_t r = nccl_api().GetUniqueId(&id);
    if (r != ncclSuccess) { bb_set_error("ncclGetUniqueId: %s", nccl_api().GetErrorString(r)); return BB_ERR_NCCL; }
    memcpy(out128, &id, 128);
    return BB_OK;
}

// Communicators outlive handles: a worker that runs several simulations on the same slab layout (forward,
// back-propagation, refocus) pays ncclCommInitRank once.  Keyed by (device, rank, nranks); id128 == NULL
// re-attaches the cached communicator of that key.
struct CommKey { int device, rank, nranks; bool operator<(const CommKey &o) const { return std::tie(device, rank, nranks) < std::tie(o.device, o.rank, o.nranks); } };
static std::map<CommKey, ncclComm_t> g_comms;
static std::mutex g_comms_mutex;

extern "C" int bb_fdtd_comm_init(bb_fdtd *h, const char *id128) {
    BB_REQUIRE(h, "null argument");
    BB_REQUIRE(h->d.nranks > 1, "comm_init needs nranks > 1");
    if (!nccl_api().ok) { bb_set_error("NCCL not available: %s", nccl_api().err.c_str()); return BB_ERR_NCCL; }
    BB_CUDA(cudaSetDevice(h->d.device));
    const CommKey key{h->d.device, h->d.rank, h->d.nranks};
    ncclComm_t old = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_comms_mutex);
        auto it = g_comms.find(key);
        if (!id128) {
            if (it == g_comms.end()) { bb_set_error("no cached communicator for rank %d of %d on device %d", key.rank, key.nranks, key.device); return BB_ERR_STATE; }
            h->comm = it->second;
            return BB_OK;
        }
        if (it != g_comms.end()) { old = it->second; g_comms.erase(it); }
    }
    // ncclCommInitRank blocks until every rank has joined: the cache lock must not be held here (ranks may be threads)
    if (old) nccl_api().CommDestroy(old);
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm = nullptr;
    ncclResult_t r = nccl_api().CommInitRank(&comm, h->d.nranks, id, h->d.rank);
    if (r != ncclSuccess) { bb_set_error("ncclCommInitRank: %s", nccl_api().GetErrorString(r)); return BB_ERR_NCCL; }
    {
        std::lock_guard<std::mutex> lock(g_comms_mutex);
        g_comms[key] = comm;
    }
    h->comm = comm;
    return BB_OK;
}

// ------------------------------------------------------------------------------------------
// NVLink halo push: descriptors of the slab for its neighbours
// ------------------------------------------------------------------------------------------
extern "C" int bb_fdtd_peer_export(bb_fdtd *h, bb_peer_info *out) {
    BB_REQUIRE(h && out, "null argument");
    BB_CUDA(cudaSetDevice(h->d.device));
    memset(out, 0, sizeof(*out));
    out->pid = (int64_t)getpid();
    out->device = h->d.device; out->nown = h->nown; out->n2 = h->p.n2; out->pitch = h->p.pitch;
    out->v_ptr = (uint64_t)h->p.V[0]; out->s_ptr = (uint64_t)h->p.S[0]; out->flag_ptr = (uint64_t)h->flags;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    BB_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)out->v_ipc, h->p.V[0]));
    BB_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)out->s_ipc, h->p.S[0]));
    BB_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)out->flag_ipc, h->flags));
    return BB_OK;
}

static int map_peer(bb_fdtd *h, const bb_peer_info *nb, int side) {
    DevParams &p = h->p;
    BB_REQUIRE(nb->n2 == p.n2 && nb->pitch == p.pitch, "neighbour slab has a different transverse geometry");
    void *v = nullptr, *s = nullptr, *f = nullptr;
    if (nb->pid == (int64_t)getpid()) {          // same process (one thread per GPU): plain peer pointers
        if (nb->device != h->d.device) {
            int can = 0;
            BB_CUDA(cudaDeviceCanAccessPeer(&can, h->d.device, nb->device));
            if (!can) { bb_set_error("device %d cannot access device %d as a peer", h->d.device, nb->device); return BB_ERR_CUDA; }
            cudaError_t e = cudaDeviceEnablePeerAccess(nb->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { bb_set_error("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)); return BB_ERR_CUDA; }
            cudaGetLastError();
        }
        v = (void *)nb->v_ptr; s = (void *)nb->s_ptr; f = (void *)nb->flag_ptr;
    } else {                                     // another process: CUDA IPC mappings
        BB_CUDA(cudaIpcOpenMemHandle(&v, *(const cudaIpcMemHandle_t *)nb->v_ipc, cudaIpcMemLazyEnablePeerAccess));
        BB_CUDA(cudaIpcOpenMemHandle(&s, *(const cudaIpcMemHandle_t *)nb->s_ipc, cudaIpcMemLazyEnablePeerAccess));
        BB_CUDA(cudaIpcOpenMemHandle(&f, *(const cudaIpcMemHandle_t *)nb->flag_ipc, cudaIpcMemLazyEnablePeerAccess));
        h->ipc_opened[3 * side] = v; h->ipc_opened[3 * side + 1] = s; h->ipc_opened[3 * side + 2] = f;
    }
    p.peerV[side] = (float *)v; p.peerS[side] = (float *)s;
    p.peer_vol[side] = (long long)(nb->nown + 4) * p.plane;
    // my planes i0, i0+1 are the lower neighbour's upper halo (its local planes nown+2, nown+3);
    // my planes i1-2, i1-1 are the upper neighbour's lower halo (its local planes 0, 1)
    p.peer_plane[side] = side == 0 ? (unsigned)(nb->nown + 2) : 0u;
    // I am the lower neighbour's upper side (its flag word 1) and the upper neighbour's lower side (word 0)
    p.flag_peer[side] = (unsigned long long *)f + (side == 0 ? 1 : 0);
    return BB_OK;
}

extern "C" int bb_fdtd_peer_attach(bb_fdtd *h, const bb_peer_info *lower, const bb_peer_info *upper) {
    BB_REQUIRE(h, "null handle");
    BB_REQUIRE(h->d.kernel_variant != 1, "the NVLink halo push needs the TMA kernels (variant 0 or 2)");
    BB_REQUIRE((lower != nullptr) == (h->d.rank > 0) && (upper != nullptr) == (h->d.rank < h->d.nranks - 1),
               "rank %d of %d needs exactly its existing neighbours", h->d.rank, h->d.nranks);
    BB_CUDA(cudaSetDevice(h->d.device));
    int rc;
    if (lower && (rc = map_peer(h, lower, 0))) return rc;
    if (upper && (rc = map_peer(h, upper, 1))) return rc;
    h->peer_mode = true;
    return BB_OK;
}

// ------------------------------------------------------------------------------------------
// time loop
// ------------------------------------------------------------------------------------------
struct Timer {
    bb_fdtd *h; bool on;
    void begin(int cat) {
        if (!on) return;
        if (h->ev_used + 2 > h->ev_pool.size()) {
            for (int n = 0; n < 2; n++) { cudaEvent_t e; cudaEventCreate(&e); h->ev_pool.push_back(e); }
        }
        h->ev_cat.push_back(cat);
        cudaEventRecord(h->ev_pool[h->ev_used], h->stream);
    }
    void end() {
        if (!on) return;
        cudaEventRecord(h->ev_pool[h->ev_used + 1], h->stream);
        h->ev_used += 2;
    }
};

// Plane ranges of the CTAs of one launch over the planes [ib, ie).  One CTA per SM is resident; chunks are as long
// as the flag table allows (64 planes), and the last one is split into pieces of halving length (>= 6 planes) so
// that the final, partially filled wave costs a few planes instead of a whole chunk.
static ChunkPlan make_chunk_plan(const bb_fdtd *h, int ib, int ie, int tile_rows = tma::TY) {
    ChunkPlan pl;
    const int nplanes = ie - ib;
    const int tiles = h->p.ntk * ((h->p.n2 + tile_rows - 1) / tile_rows);
    int chunk;
    if (h->chunk_override > 0) chunk = h->chunk_override;
    else {
        // long chunks amortise the pipeline fill of a CTA (measured on CTX-500: 64-plane chunks 31.7, 30-plane 30.7,
        // 16-plane 28.9 Gcell/s); small grids still get at least two CTAs per SM
        const int nch = std::max(1, (2 * 148 + tiles - 1) / tiles);
        chunk = std::max((nplanes + nch - 1) / nch, std::min(nplanes, 8));
    }
    chunk = std::max(1, std::min(chunk, (int)tma::MAXCHUNK));
    while ((nplanes + chunk - 1) / chunk + 6 > BB_MAX_CHUNKS) chunk++;    // only above 2688 planes per GPU
    pl.n = 0;
    int at = ib;
    auto push = [&](int len) { pl.start[pl.n] = at; at += len; pl.end[pl.n++] = at; };
    while (ie - at > chunk) push(chunk);
    int rest = ie - at;                        // 1 .. chunk planes left: halve it down
    while (rest >= 12 && h->chunk_tail) { const int len = (rest + 1) / 2; push(len); rest -= len; }
    if (rest > 0) push(rest);
    // NVLink halo push: the piece(s) holding the last two planes go first, so that the upper neighbour has their
    // results early in the next half-step (the first chunk, which serves the lower neighbour, is early anyway)
    if (h->peer_mode && h->p.peerS[1] && ie == h->p.i1 && pl.n > 2) {
        const int move = (pl.end[pl.n - 1] - pl.start[pl.n - 1] >= 2) ? 1 : 2;
        for (int m = 0; m < move; m++) {
            const int s0 = pl.start[pl.n - 1], e0 = pl.end[pl.n - 1];
            for (int z = pl.n - 1; z > 0; z--) { pl.start[z] = pl.start[z - 1]; pl.end[z] = pl.end[z - 1]; }
            pl.start[0] = s0; pl.end[0] = e0;
        }
    }
    return pl;
}

static unsigned long long half_step_seq(const bb_fdtd *h, bool stress) {
    return ((unsigned long long)h->epoch << 32) | (unsigned long long)(2 * h->step + (stress ? 0 : 1) + 1);
}

template <typename K>
static int set_smem(K kernel, int bytes) {
    BB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return BB_OK;
}

template <typename LT>
static int prepare_kernels() {
    int rc;
    if ((rc = set_smem(tma::stress_tma<LT, 0, false>, tma::SMEM_BYTES))) return rc;
    if ((rc = set_smem(tma::stress_tma<LT, 1, false>, tma::SMEM_BYTES))) return rc;
    if ((rc = set_smem(tma::stress_tma<LT, 2, false>, tma::SMEM_BYTES))) return rc;
    if ((rc = set_smem(tma::particle_tma<LT, 0, false>, tma::SMEM_BYTES))) return rc;
    if ((rc = set_smem(tma::particle_tma<LT, 2, false>, tma::SMEM_BYTES))) return rc;
    if ((rc = set_smem(tma::stress_tma<LT, 0, true>, tma::SMEM_BYTES))) return rc;
    if ((rc = set_smem(tma::stress_tma<LT, 1, true>, tma::SMEM_BYTES))) return rc;
    if ((rc = set_smem(tma::stress_tma<LT, 2, true>, tma::SMEM_BYTES))) return rc;
    if ((rc = set_smem(tma::particle_tma<LT, 0, true>, tma::SMEM_BYTES))) return rc;
    if ((rc = set_smem(tma::particle_tma<LT, 2, true>, tma::SMEM_BYTES))) return rc;
    if ((rc = set_smem(tma::stress_tma2<LT, 0>, tma::SMEM_BYTES))) return rc;
    if ((rc = set_smem(tma::stress_tma2<LT, 1>, tma::SMEM_BYTES))) return rc;
    if ((rc = set_smem(tma::stress_tma2<LT, 2>, tma::SMEM_BYTES))) return rc;
    if ((rc = set_smem(tma::particle_tma2<LT, 0, 8>, tma::SMEM_BYTES))) return rc;
    if ((rc = set_smem(tma::particle_tma2<LT, 2, 8>, tma::SMEM_BYTES))) return rc;
    if ((rc = set_smem(tma::particle_tma2<LT, 0, 16>, tma::SMEM_BYTES))) return rc;
    if ((rc = set_smem(tma::particle_tma2<LT, 2, 16>, tma::SMEM_BYTES))) return rc;
    return BB_OK;
}

// one half-step over the owned planes [ib, ie): a single fused launch (interior + PML shell)
template <typename LT>
static int launch_half_step(bb_fdtd *h, bool stress, int acc_mode, int ib, int ie, Timer &tm, const ChunkPlan *given = nullptr,
                            bool publish_in_kernel = true, int src_step = -1) {
    if (ie <= ib) return BB_OK;
    const DevParams &p = h->p;
    tm.begin(stress ? CAT_STRESS : CAT_PARTICLE);
    if (h->d.kernel_variant == 1) {
        const dim3 blk(64, 4, 1), grid((p.n3 + 63) / 64, (p.n2 + 3) / 4, ie - ib);
        if (stress) {
            if (acc_mode) direct::stress_direct<LT, true><<<grid, blk, 0, h->stream>>>(p, ib);
            else direct::stress_direct<LT, false><<<grid, blk, 0, h->stream>>>(p, ib);
        } else {
            if (acc_mode) direct::particle_direct<LT, true><<<grid, blk, 0, h->stream>>>(p, ib);
            else direct::particle_direct<LT, false><<<grid, blk, 0, h->stream>>>(p, ib);
        }
    } else {
        const int variant = h->d.kernel_variant;
        const bool p16 = !stress && variant == 3;                 // particle half-step on 16-row tiles, two cells per thread
        const ChunkPlan plan = given ? *given : make_chunk_plan(h, ib, ie, p16 ? 16 : tma::TY);
        DevParams p = h->p;       // per-launch copy: the sequence number of this half-step for the NVLink halo push
        p.seq = half_step_seq(h, stress);
        p.publish = publish_in_kernel ? 1 : 0;
        p.bsrc_map = nullptr;
        if (src_step >= 0 && h->bsrc_map && h->nsrc_boundary > 0) {     // this launch injects the boundary-plane sources of time step src_step
            p.bsrc_map = h->bsrc_map; p.bsrc_row = h->src_row;
            p.bsrc_o[0] = h->src_o[0]; p.bsrc_o[1] = h->src_o[1]; p.bsrc_o[2] = h->src_o[2];
            p.sf_row = h->tone_ac ? nullptr : h->srcfun + (size_t)src_step * h->d.nsrc;
            p.tone_ac = h->tone_ac; p.tone_as = h->tone_as;
            p.env_sin = h->tone_ac ? h->tone_es[src_step] : 0.f; p.env_cos = h->tone_ac ? h->tone_ec[src_step] : 0.f;
            p.src_hard = (h->d.type_source & 1);
        }
        if (h->dbg_kernel >= 0 && h->dbg_kernel != (stress ? 0 : 1)) p.dbg = nullptr;   // BB_CTA_TIMING=stress|particle
        const int sm = tma::SMEM_BYTES;
        if (p16) {
            const dim3 grid(p.ntk, (p.n2 + 15) / 16, plan.n), blk(tma::PT<16>::NTB, 1, 1);
            if (acc_mode) tma::particle_tma2<LT, 2, 16><<<grid, blk, sm, h->stream>>>(h->pmaps16, p, plan);
            else tma::particle_tma2<LT, 0, 16><<<grid, blk, sm, h->stream>>>(h->pmaps16, p, plan);
        } else if (variant == 2) {      // two cells per thread on the 8-row tiles (fdtd_tma2.cuh)
            const dim3 grid(p.ntk, p.ntj, plan.n), blk(tma::NTB2, 1, 1);
            if (stress) {
                if (acc_mode == 1) tma::stress_tma2<LT, 1><<<grid, blk, sm, h->stream>>>(h->smaps, p, plan);
                else if (acc_mode == 2) tma::stress_tma2<LT, 2><<<grid, blk, sm, h->stream>>>(h->smaps, p, plan);
                else tma::stress_tma2<LT, 0><<<grid, blk, sm, h->stream>>>(h->smaps, p, plan);
            } else {
                if (acc_mode) tma::particle_tma2<LT, 2, 8><<<grid, blk, sm, h->stream>>>(h->pmaps, p, plan);
                else tma::particle_tma2<LT, 0, 8><<<grid, blk, sm, h->stream>>>(h->pmaps, p, plan);
            }
        } else {
            const dim3 grid(p.ntk, p.ntj, plan.n), blk(tma::NTB, 1, 1);
            // the kernels are compiled twice: for a slab with a neighbour on some side (halo wait, push, publish) and without
            const bool peer = p.peerV[0] || p.peerV[1] || p.peerS[0] || p.peerS[1];
            if (stress) {
                auto k = acc_mode == 1 ? (peer ? tma::stress_tma<LT, 1, true> : tma::stress_tma<LT, 1, false>)
                       : acc_mode == 2 ? (peer ? tma::stress_tma<LT, 2, true> : tma::stress_tma<LT, 2, false>)
                                       : (peer ? tma::stress_tma<LT, 0, true> : tma::stress_tma<LT, 0, false>);
                k<<<grid, blk, sm, h->stream>>>(h->smaps, p, plan);
            } else {
                auto k = acc_mode ? (peer ? tma::particle_tma<LT, 2, true> : tma::particle_tma<LT, 2, false>)
                                  : (peer ? tma::particle_tma<LT, 0, true> : tma::particle_tma<LT, 0, false>);
                k<<<grid, blk, sm, h->stream>>>(h->pmaps, p, plan);
            }
        }
    }
    tm.end();
    BB_CUDA(cudaGetLastError());
    return BB_OK;
}

static int launch_sources(bb_fdtd *h, int n, int64_t first, int64_t count, Timer &tm) {
    if (count <= 0 || n >= h->d.nt_src || !(h->srcfun || h->tone_ac)) return BB_OK;
    tm.begin(CAT_OTHER);
    DevParams sp = h->p;
    // NVLink halo push: the source cells of the pushed planes are injected by the half-step kernels (DevParams::bsrc_map);
    // the cells handled here lie in no pushed plane, so this kernel neither stores to a neighbour nor fences system-wide
    // (one fence.sys per source thread made this 15 us kernel cost 45 us of every time step at N = 2)
    if (h->peer_mode) { sp.peerV[0] = sp.peerV[1] = sp.peerS[0] = sp.peerS[1] = nullptr; }
    source_kernel<<<(unsigned)((count + 127) / 128), 128, 0, h->stream>>>(sp, h->d.type_source, count, h->src_cell + first, h->src_row + first,
                                                                         h->src_o[0] + first, h->src_o[1] + first, h->src_o[2] + first,
                                                                         h->tone_ac ? nullptr : h->srcfun + (size_t)n * h->d.nsrc,
                                                                         h->tone_ac, h->tone_as, h->tone_ac ? h->tone_es[n] : 0.f,
                                                                         h->tone_ac ? h->tone_ec[n] : 0.f);
    tm.end();
    BB_CUDA(cudaGetLastError());
    return BB_OK;
}

// exchange two planes of three fields with both slab neighbours (NCCL send/recv on comm_stream)
static int halo_exchange(bb_fdtd *h, float *const f[3]) {
    const DevParams &p = h->p;
    const size_t cnt = (size_t)2 * p.plane;
    const int r = h->d.rank, nr = h->d.nranks;
    NcclApi &api = nccl_api();
    ncclResult_t rc = api.GroupStart();
    for (int c = 0; c < 3 && rc == ncclSuccess; c++) {
        if (r > 0) {
            rc = api.Send(f[c] + 2 * p.plane, cnt, ncclFloat, r - 1, h->comm, h->comm_stream);
            if (rc == ncclSuccess) rc = api.Recv(f[c], cnt, ncclFloat, r - 1, h->comm, h->comm_stream);
        }
        if (r < nr - 1 && rc == ncclSuccess) {
            rc = api.Send(f[c] + (size_t)h->nown * p.plane, cnt, ncclFloat, r + 1, h->comm, h->comm_stream);
            if (rc == ncclSuccess) rc = api.Recv(f[c] + (size_t)(h->nown + 2) * p.plane, cnt, ncclFloat, r + 1, h->comm, h->comm_stream);
        }
    }
    ncclResult_t rc2 = api.GroupEnd();
    if (rc == ncclSuccess) rc = rc2;
    if (rc != ncclSuccess) { bb_set_error("NCCL halo exchange: %s", api.GetErrorString(rc)); return BB_ERR_NCCL; }
    return BB_OK;
}

template <typename LT>
static int half_step(bb_fdtd *h, bool stress, int n, int acc, Timer &tm) {
    const DevParams &p = h->p;
    const bool src_here = stress ? (h->d.type_source >= 2) : (h->d.type_source < 2);
    int rc;
    if (h->peer_mode) {
        // NVLink halo push: one launch; the boundary CTAs store into the neighbours' halo planes, inject the sources that sit
        // in those planes themselves and publish the half-step as soon as the planes are complete.  The source kernel behind
        // it handles the other source cells (measured before this: publishing behind the source kernel cost 5 % at N = 2,
        // profiles/r2_scaling_experiments.txt).
        static const bool nosrc = getenv("BB_EXPERIMENT_NOSRC") != nullptr;     // scaling experiments only
        const bool src_now = !nosrc && src_here && h->nsrc_cells > 0 && n < h->d.nt_src && (h->srcfun || h->tone_ac);
        if ((rc = launch_half_step<LT>(h, stress, acc, p.i0, p.i1, tm, nullptr, true, src_now ? n : -1))) return rc;
        if (src_now && (rc = launch_sources(h, n, h->nsrc_boundary, h->nsrc_cells - h->nsrc_boundary, tm))) return rc;
        return BB_OK;
    }
    if (h->d.nranks > 1) {
        // inputs of this half-step: halos sent during the previous half-step
        BB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_halo, 0));
        // the two boundary plane pairs in one launch (variant 0), so their results can leave while the interior runs
        if (h->d.kernel_variant != 1) {
            ChunkPlan edge;
            edge.n = 2;
            edge.start[0] = p.i0; edge.end[0] = p.i0 + 2; edge.start[1] = p.i1 - 2; edge.end[1] = p.i1;
            if ((rc = launch_half_step<LT>(h, stress, acc, p.i0, p.i1, tm, &edge))) return rc;
        } else {
            if ((rc = launch_half_step<LT>(h, stress, acc, p.i0, p.i0 + 2, tm))) return rc;
            if ((rc = launch_half_step<LT>(h, stress, acc, p.i1 - 2, p.i1, tm))) return rc;
        }
        if (src_here && (rc = launch_sources(h, n, 0, h->nsrc_boundary, tm))) return rc;
        BB_CUDA(cudaEventRecord(h->ev_boundary, h->stream));
        BB_CUDA(cudaStreamWaitEvent(h->comm_stream, h->ev_boundary, 0));
        float *sf[3] = { p.S[0], p.S[3], p.S[4] };
        float *vf[3] = { p.V[0], p.V[1], p.V[2] };
        if ((rc = halo_exchange(h, stress ? sf : vf))) return rc;
        BB_CUDA(cudaEventRecord(h->ev_halo, h->comm_stream));
        if ((rc = launch_half_step<LT>(h, stress, acc, p.i0 + 2, p.i1 - 2, tm))) return rc;
        if (src_here && (rc = launch_sources(h, n, h->nsrc_boundary, h->nsrc_cells - h->nsrc_boundary, tm))) return rc;
    } else {
        if ((rc = launch_half_step<LT>(h, stress, acc, p.i0, p.i1, tm))) return rc;
        if (src_here && (rc = launch_sources(h, n, 0, h->nsrc_cells, tm))) return rc;
    }
    return BB_OK;
}

template <typename LT>
static int run_steps(bb_fdtd *h, int64_t nsteps, Timer &tm) {
    const bb_fdtd_desc &d = h->d;
    const int n0 = d.sensor_start * d.sensor_subsampling;
    const unsigned stress_maps = d.sel_maps_rms & 0x7F0u, part_maps = d.sel_maps_rms & 0xFu;
    int rc;
    for (int64_t t = 0; t < nsteps; t++) {
        const int n = (int)h->step;
        if ((rc = ensure_source_samples(h, n))) return rc;
        const bool window = d.sel_rms_peak != 0 && n >= n0;
        const bool only_p_rms = d.sel_rms_peak == 1 && d.sel_maps_rms == (1u << BB_MAP_PRESSURE);
        if ((rc = half_step<LT>(h, true, n, (window && stress_maps) ? (only_p_rms ? 1 : 2) : 0, tm))) return rc;
        if ((rc = half_step<LT>(h, false, n, (window && part_maps) ? 2 : 0, tm))) return rc;
        if (h->nsensors && h->n_sensor_maps && n % d.sensor_subsampling == 0 && n / d.sensor_subsampling >= d.sensor_start) {
            const long long sample = n / d.sensor_subsampling - d.sensor_start;
            tm.begin(CAT_OTHER);
            sensor_kernel<LT><<<(unsigned)((h->nsensors + 255) / 256), 256, 0, h->stream>>>(h->p, d.sel_maps_sensor, h->nsensors, h->sensor_cell,
                                                                                           h->sensor_out, h->nsamples, sample);
            tm.end();
            BB_CUDA(cudaGetLastError());
        }
        h->step++;
    }
    return BB_OK;
}

extern "C" int bb_fdtd_run(bb_fdtd *h, int64_t nsteps, int profile) {
    BB_REQUIRE(h, "null handle");
    if (!h->materials_set || !h->maps_set) { bb_set_error("set_materials / set_maps must be called before run"); return BB_ERR_STATE; }
    // timing experiments only (profiles/run_slab_cost.py): run one slab of a decomposed grid alone, halos left as uploaded
    if (h->d.nranks > 1 && !h->comm && !h->peer_mode && getenv("BB_EXPERIMENT_NOHALO")) h->peer_mode = true;
    if (h->d.nranks > 1 && !h->comm && !h->peer_mode) { bb_set_error("multi-rank handle without comm_init / peer_attach"); return BB_ERR_STATE; }
    BB_CUDA(cudaSetDevice(h->d.device));
    if (nsteps < 0 || h->step + nsteps > h->d.steps) nsteps = h->d.steps - h->step;
    if (!h->prepared) {
        const dim3 blk(tma::TX, tma::TY, 1), grid(h->p.ntk, h->p.ntj, h->p.nloc);
        int rcp;
        if (h->label_bytes == 1) { tma::flags_kernel<uint8_t><<<grid, blk, 0, h->stream>>>(h->p, (unsigned char *)h->p.flags); rcp = prepare_kernels<uint8_t>(); }
        else { tma::flags_kernel<uint16_t><<<grid, blk, 0, h->stream>>>(h->p, (unsigned char *)h->p.flags); rcp = prepare_kernels<uint16_t>(); }
        BB_CUDA(cudaGetLastError());
        if (rcp) return rcp;
        h->prepared = true;
    }
    Timer tm{h, profile != 0};
    h->ev_used = 0;
    h->ev_cat.clear();
    if (h->d.nranks > 1 && !h->peer_mode) BB_CUDA(cudaEventRecord(h->ev_halo, h->comm_stream));
    BB_CUDA(cudaEventRecord(h->ev_run0, h->stream));
    int rc = h->label_bytes == 1 ? run_steps<uint8_t>(h, nsteps, tm) : run_steps<uint16_t>(h, nsteps, tm);
    if (rc) return rc;
    if (h->d.nranks > 1 && !h->peer_mode) BB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_halo, 0));
    BB_CUDA(cudaEventRecord(h->ev_run1, h->stream));
    int herr = 0;
    if (h->peer_mode) BB_CUDA(cudaMemcpyAsync(&herr, h->p.err, 4, cudaMemcpyDeviceToHost, h->stream));
    BB_CUDA(cudaStreamSynchronize(h->stream));
    BB_CUDA(cudaGetLastError());
    if (herr) {
        BB_CUDA(cudaMemsetAsync(h->p.err, 0, 4, h->stream));
        bb_set_error("rank %d: the halo planes of the %s neighbour did not arrive within BB_PEER_TIMEOUT_S; results are invalid", h->d.rank, herr == 1 ? "lower" : "upper");
        return BB_ERR_STATE;
    }
    float ms = 0;
    BB_CUDA(cudaEventElapsedTime(&ms, h->ev_run0, h->ev_run1));
    bb_fdtd_stats &st = h->stats;
    st.run_ms = ms;
    st.stress_ms = st.particle_ms = st.pml_ms = st.other_ms = 0;
    st.stress_launches = st.particle_launches = st.pml_launches = st.other_launches = 0;
    for (size_t e = 0; e < h->ev_cat.size(); e++) {
        float t = 0;
        cudaEventElapsedTime(&t, h->ev_pool[2 * e], h->ev_pool[2 * e + 1]);
        switch (h->ev_cat[e]) {
        case CAT_STRESS: st.stress_ms += t; st.stress_launches++; break;
        case CAT_PARTICLE: st.particle_ms += t; st.particle_launches++; break;
        case CAT_PML: st.pml_ms += t; st.pml_launches++; break;
        default: st.other_ms += t; st.other_launches++; break;
        }
    }
    st.steps_done = h->step;
    return BB_OK;
}

extern "C" int bb_fdtd_reset(bb_fdtd *h) {
    BB_REQUIRE(h, "null handle");
    BB_CUDA(cudaSetDevice(h->d.device));
    const DevParams &p = h->p;
    const size_t vol = (size_t)p.nloc * p.plane * 4;
    BB_CUDA(cudaMemsetAsync(p.V[0], 0, 3 * vol, h->stream));
    BB_CUDA(cudaMemsetAsync(p.S[0], 0, 6 * vol, h->stream));
    BB_CUDA(cudaMemsetAsync(p.R[0], 0, 6 * vol, h->stream));
    BB_CUDA(cudaMemsetAsync(p.Pr, 0, vol, h->stream));
    BB_CUDA(cudaMemsetAsync(p.XP[0], 0, BB_NPART * h->xp_floats * 4, h->stream));
    BB_CUDA(cudaMemsetAsync(p.YP[0], 0, BB_NPART * h->yp_floats * 4, h->stream));
    BB_CUDA(cudaMemsetAsync(p.ZP[0], 0, BB_NPART * h->zp_floats * 4, h->stream));
    if (p.acc_rms) BB_CUDA(cudaMemsetAsync(p.acc_rms, 0, (size_t)h->n_acc_maps * p.acc_stride * 4, h->stream));
    if (p.acc_peak) BB_CUDA(cudaMemsetAsync(p.acc_peak, 0, (size_t)h->n_acc_maps * p.acc_stride * 4, h->stream));
    if (h->sensor_out) BB_CUDA(cudaMemsetAsync(h->sensor_out, 0, (size_t)h->n_sensor_maps * h->nsamples * h->nsensors * 4, h->stream));
    BB_CUDA(cudaStreamSynchronize(h->stream));
    h->step = 0;
    h->epoch++;                   // sequence numbers of the NVLink halo push never repeat: flags of an earlier epoch are just smaller
    return BB_OK;
}

extern "C" int bb_fdtd_get_map(bb_fdtd *h, int which, int map_id, float *out) {
    BB_REQUIRE(h && out, "null argument");
    BB_REQUIRE(map_id >= 0 && map_id < BB_MAP_COUNT && which >= 0 && which <= 2, "bad map selector");
    BB_CUDA(cudaSetDevice(h->d.device));
    const DevParams &p = h->p;
    const long long nrows = (long long)h->nown * p.n2;
    const float *src = nullptr;
    int mode = 0;
    float inv_nacc = 0.f;
    if (which == 2) {
        BB_REQUIRE(map_id != BB_MAP_ALLV, "no last map for ALLV");
        if (map_id != BB_MAP_PRESSURE) src = (map_id <= BB_MAP_VZ ? p.V[map_id - BB_MAP_VX] : p.S[map_id - BB_MAP_SXX]) + 2 * p.plane;
    } else {
        const float *base = which == 0 ? p.acc_rms : p.acc_peak;
        if (!base || !(h->d.sel_maps_rms & (1u << map_id))) { bb_set_error("map %d was not selected", map_id); return BB_ERR_ARG; }
        const int slot = popcount32(h->d.sel_maps_rms & ((1u << map_id) - 1u));
        const int n0 = h->d.sensor_start * h->d.sensor_subsampling;
        const long long nacc = std::max<long long>(1, h->d.steps - n0);
        mode = which == 0 ? 1 : (map_id == BB_MAP_ALLV ? 2 : 0);
        inv_nacc = 1.0f / (float)nacc;
        src = base + (size_t)slot * p.acc_stride;
    }
    // dense rows leave through the staging slots, a group of whole rows at a time (no per-call device allocation)
    const size_t row_bytes = (size_t)p.n3 * 4;
    const long long rows_per = std::max<long long>(1, (long long)(BB_STAGE_BYTES / row_bytes));
    BB_REQUIRE(row_bytes <= BB_STAGE_BYTES, "a row of %d values does not fit the staging buffer", p.n3);
    return staged_download(h, out, (size_t)nrows * row_bytes, (size_t)rows_per * row_bytes, [&](void *dev, size_t off, size_t n) {
        const long long r0 = (long long)(off / row_bytes), nr = (long long)(n / row_bytes);
        const unsigned grid = (unsigned)((nr * p.n3 + 255) / 256);
        if (!src) {
            if (h->label_bytes == 1) finalize_pressure_kernel<uint8_t><<<grid, 256, 0, h->stream>>>(p, (float *)dev, r0, nr);
            else finalize_pressure_kernel<uint16_t><<<grid, 256, 0, h->stream>>>(p, (float *)dev, r0, nr);
        } else {
            finalize_map_kernel<<<grid, 256, 0, h->stream>>>(src + r0 * p.pitch, (float *)dev, nr, p.n3, p.pitch, mode, inv_nacc);
        }
        return BB_OK;
    });
}

extern "C" int bb_fdtd_get_sensors(bb_fdtd *h, int map_id, float *out) {
    BB_REQUIRE(h && out, "null argument");
    BB_REQUIRE(map_id >= 0 && map_id < BB_MAP_COUNT && (h->d.sel_maps_sensor & (1u << map_id)), "sensor map %d not selected", map_id);
    BB_CUDA(cudaSetDevice(h->d.device));
    if (h->nsensors == 0 || h->nsamples == 0) return BB_OK;
    const int slot = popcount32(h->d.sel_maps_sensor & ((1u << map_id) - 1u));
    const int nsam = (int)h->nsamples;
    const size_t smem = (size_t)nsam * (BB_ST_SENSORS + 1) * 4;
    BB_REQUIRE(smem <= 200u * 1024u, "%d samples per sensor exceed the transpose tile", nsam);
    if (smem > 48u * 1024u) BB_CUDA(cudaFuncSetAttribute(sensor_transpose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const float *in = h->sensor_out + (size_t)slot * h->nsensors * h->nsamples;
    const size_t row_bytes = (size_t)nsam * 4;
    const long long per = std::max<long long>(BB_ST_SENSORS, (long long)(BB_STAGE_BYTES / row_bytes) / BB_ST_SENSORS * BB_ST_SENSORS);
    BB_REQUIRE((size_t)per * row_bytes <= BB_STAGE_BYTES, "%d samples per sensor do not fit the staging buffer", nsam);
    return staged_download(h, out, (size_t)h->nsensors * row_bytes, (size_t)per * row_bytes, [&](void *dev, size_t off, size_t n) {
        const long long s0 = (long long)(off / row_bytes), ns = (long long)(n / row_bytes);
        sensor_transpose_kernel<<<(unsigned)((ns + BB_ST_SENSORS - 1) / BB_ST_SENSORS), BB_ST_SENSORS, smem, h->stream>>>(in, (float *)dev, h->nsensors, nsam, s0, ns);
        return BB_OK;
    });
}

extern "C" int bb_fdtd_get_sensors_runs(bb_fdtd *h, int map_id, float *table, int64_t table_rows, const int64_t *dst_row,
                                        const int64_t *src_row, const int64_t *nrows, int64_t nruns, int *done) {
    BB_REQUIRE(h && table && done, "null argument");
    BB_REQUIRE(map_id >= 0 && map_id < BB_MAP_COUNT && (h->d.sel_maps_sensor & (1u << map_id)), "sensor map %d not selected", map_id);
    BB_REQUIRE(nruns >= 0 && (nruns == 0 || (dst_row && src_row && nrows)), "bad run tables");
    *done = 0;
    BB_CUDA(cudaSetDevice(h->d.device));
    if (!host_is_pinned(table)) return BB_OK;                // the caller falls back to bb_fdtd_get_sensors + bb_host_scatter_runs
    *done = 1;
    if (h->nsensors == 0 || h->nsamples == 0 || nruns == 0) return BB_OK;
    int64_t at = 0;                                          // the runs must tile the slab's rows in order and stay inside the table
    for (int64_t u = 0; u < nruns; u++) {
        if (nrows[u] <= 0 || src_row[u] != at || dst_row[u] < 0 || dst_row[u] + nrows[u] > table_rows) {
            bb_set_error("run %lld (rows %lld..+%lld -> %lld) does not continue the slab's rows or leaves the table of %lld rows",
                         (long long)u, (long long)src_row[u], (long long)nrows[u], (long long)dst_row[u], (long long)table_rows);
            return BB_ERR_ARG;
        }
        at += nrows[u];
    }
    BB_REQUIRE(at == h->nsensors, "the runs cover %lld rows, the slab has %lld sensors", (long long)at, (long long)h->nsensors);
    const int slot = popcount32(h->d.sel_maps_sensor & ((1u << map_id) - 1u));
    const int nsam = (int)h->nsamples;
    const size_t smem = (size_t)nsam * (BB_ST_SENSORS + 1) * 4;
    BB_REQUIRE(smem <= 200u * 1024u, "%d samples per sensor exceed the transpose tile", nsam);
    if (smem > 40u * 1024u) BB_CUDA(cudaFuncSetAttribute(sensor_transpose_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DevTmp tsrc(h->d.device), tdst(h->d.device);
    cudaError_t e = tsrc.alloc((size_t)nruns * 8);
    if (e == cudaSuccess) e = tdst.alloc((size_t)nruns * 8);
    if (e == cudaSuccess) e = cudaMemcpyAsync(tsrc.as<long long>(), src_row, (size_t)nruns * 8, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(tdst.as<long long>(), dst_row, (size_t)nruns * 8, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
        const float *in = h->sensor_out + (size_t)slot * h->nsensors * h->nsamples;
        float *dtab = nullptr;
        e = cudaHostGetDevicePointer((void **)&dtab, table, 0);
        if (e == cudaSuccess) {
            sensor_transpose_scatter_kernel<<<(unsigned)((h->nsensors + BB_ST_SENSORS - 1) / BB_ST_SENSORS), BB_ST_SENSORS, smem, h->stream>>>(
                in, dtab, h->nsensors, nsam, tsrc.as<long long>(), tdst.as<long long>(), nruns);
            e = cudaGetLastError();
        }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) { bb_set_error("get_sensors_runs: %s", cudaGetErrorString(e)); cudaGetLastError(); return BB_ERR_CUDA; }
    return BB_OK;
}

extern "C" int bb_fdtd_get_phase_data(bb_fdtd *h, int map_id, int bin, int nsamples_used, float scale,
                                      float *fourier_reim, float *phase, float *peak) {
    BB_REQUIRE(h && fourier_reim, "null argument");
    BB_REQUIRE(map_id >= 0 && map_id < BB_MAP_COUNT && (h->d.sel_maps_sensor & (1u << map_id)), "sensor map %d not selected", map_id);
    BB_REQUIRE(nsamples_used >= 1 && nsamples_used <= h->nsamples, "nsamples_used %d outside [1, %lld]", nsamples_used, (long long)h->nsamples);
    BB_REQUIRE(bin >= 0 && bin < nsamples_used, "DFT bin %d outside [0, %d)", bin, nsamples_used);
    BB_CUDA(cudaSetDevice(h->d.device));
    const size_t cells = (size_t)h->nown * h->p.n2 * h->p.n3;
    const int slot = popcount32(h->d.sel_maps_sensor & ((1u << map_id) - 1u));
    std::vector<float2> tw(nsamples_used);
    for (int n = 0; n < nsamples_used; n++) {
        const double a = -2.0 * M_PI * (double)(((long long)bin * n) % nsamples_used) / (double)nsamples_used;
        tw[n] = make_float2((float)cos(a), (float)sin(a));
    }
    DevTmp ttw(h->d.device), tfou(h->d.device), tph(h->d.device), tpk(h->d.device);      // cached between calls, freed on every path
    cudaError_t e = ttw.alloc(tw.size() * sizeof(float2));
    if (e == cudaSuccess) e = tfou.alloc(cells * sizeof(float2));
    if (e == cudaSuccess && phase) e = tph.alloc(cells * 4);
    if (e == cudaSuccess && peak) e = tpk.alloc(cells * 4);
    float2 *dtw = ttw.as<float2>(), *dfou = tfou.as<float2>();
    float *dph = phase ? tph.as<float>() : nullptr, *dpk = peak ? tpk.as<float>() : nullptr;
    if (e == cudaSuccess) e = cudaMemcpyAsync(dtw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(dfou, 0, cells * sizeof(float2), h->stream);
    if (e == cudaSuccess && dph) e = cudaMemsetAsync(dph, 0, cells * 4, h->stream);
    if (e == cudaSuccess && dpk) e = cudaMemsetAsync(dpk, 0, cells * 4, h->stream);
    if (e == cudaSuccess && h->nsensors > 0) {
        phase_data_kernel<<<(unsigned)((h->nsensors + 255) / 256), 256, 0, h->stream>>>(
            h->p, h->sensor_out + (size_t)slot * h->nsensors * h->nsamples, h->sensor_cell, h->nsensors, nsamples_used, dtw, scale, dfou, dph, dpk);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(fourier_reim, dfou, cells * sizeof(float2), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && dph) e = cudaMemcpyAsync(phase, dph, cells * 4, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && dpk) e = cudaMemcpyAsync(peak, dpk, cells * 4, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) { bb_set_error("bb_fdtd_get_phase_data: %s", cudaGetErrorString(e)); return BB_ERR_CUDA; }
    return BB_OK;
}

// profiling aid: the per-CTA records of the most recent half-step launch (needs BB_CTA_TIMING=1 at create); out = 4 x n uint64
extern "C" int bb_fdtd_debug_cta_times(bb_fdtd *h, unsigned long long *out, int64_t n) {
    BB_REQUIRE(h && out && n > 0 && n <= 65536, "bad argument");
    BB_REQUIRE(h->p.dbg, "create the handle with BB_CTA_TIMING=1 in the environment");
    BB_CUDA(cudaSetDevice(h->d.device));
    BB_CUDA(cudaMemcpy(out, h->p.dbg, (size_t)n * 32, cudaMemcpyDeviceToHost));
    return BB_OK;
}

extern "C" int bb_fdtd_get_stats(bb_fdtd *h, bb_fdtd_stats *out) {
    BB_REQUIRE(h && out, "null argument");
    *out = h->stats;
    out->cells_local = (int64_t)h->nown * h->p.n2 * h->p.n3;
    out->device_bytes = h->device_bytes;
    out->nsamples = h->nsamples;
    out->steps_done = h->step;
    return BB_OK;
}
