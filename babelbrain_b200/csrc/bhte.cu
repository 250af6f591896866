// Pennes bio-heat transfer equation + CEM43 thermal dose on the FDTD grid; replaces
// BabelViscoFDTD.tools.RayleighAndBHTE.BHTE / BHTEMultiplePressureFields
// (ThermalModeling/CalculateTemperatureEffects.py:14, :365-395, :406, :439, :960-990).
//
// Explicit scheme, one kernel launch per time step, temperature and dose ping-pong between two buffers:
//   T'(c) = T(c) + bh[m] (sum of the six neighbours - 6 T(c)) + perf[m] (T_core - T(c)) + Q(c) [while the beam is on]
//   dose' = dose + CEM43 increment of the step (R = 0.5 above 43 C, 0.25 below; exact integral of R^(43 - T) along the
//           linear temperature ramp of the step, split at 43 C when the ramp crosses it), in seconds
// bh[m] = kappa dt / (rho c_t h^2), perf[m] = w_b/60 1e-6 rho_b c_b dt / c_t, Q = p^2 dt Absorption (1 - exp(-2 h alpha)) /
// (2 rho^2 c h c_t) x DutyCycle are formed on the host (babelbrain_b200/thermal.py).  Faces of the volume keep their value.
// HBM-bound: 4 (T) + 4 (T') + 4 + 4 (dose) + 4 (Q) + 2 (label) = 22 bytes per cell and step; the six neighbour reads hit L1/L2.
#include <mutex>
#include "common.h"

namespace {
constexpr int BX = 64, BY = 4, BZ = 2;      // threads: k (contiguous) x j x i

__device__ __forceinline__ float cem43_step(float t0, float t1, float dt) {
    const float r2 = t1 >= 43.0f ? 0.5f : 0.25f;
    if (fabsf(t1 - t0) < 1.0e-4f) return dt * powf(r2, 43.0f - t1);
    const float r1 = t0 >= 43.0f ? 0.5f : 0.25f;
    if (r1 == r2) return (powf(r2, 43.0f - t1) - powf(r1, 43.0f - t0)) / (-(t1 - t0) / dt * logf(r1));
    const float dtp = dt * (43.0f - t0) / (t1 - t0);
    return (1.0f - powf(r1, 43.0f - t0)) / (-(43.0f - t0) / dtp * logf(r1))
         + (powf(r2, 43.0f - t1) - 1.0f) / ((43.0f - t1) / (dt - dtp) * logf(r2));
}

__global__ void __launch_bounds__(BX * BY * BZ) bhte_kernel(int n1, int n2, int n3, const float *__restrict__ tin, float *__restrict__ tout,
                                                            const float *__restrict__ din, float *__restrict__ dout,
                                                            const unsigned short *__restrict__ lab, const float *__restrict__ bh,
                                                            const float *__restrict__ perf, const float *__restrict__ q, float core, float dt,
                                                            int sel_j, float *__restrict__ slice, long long slice_steps, long long slice_col,
                                                            const unsigned *__restrict__ points, float *__restrict__ temp_points,
                                                            long long total_steps, long long step) {
    const int k = blockIdx.x * BX + threadIdx.x, j = blockIdx.y * BY + threadIdx.y, i = blockIdx.z * BZ + threadIdx.z;
    if (k >= n3 || j >= n2 || i >= n1) return;
    const long long s2 = n3, s1 = (long long)n2 * n3, c = (long long)i * s1 + (long long)j * s2 + k;
    const float t0 = tin[c];
    float t1 = t0, d1 = din[c];
    if (i > 0 && i < n1 - 1 && j > 0 && j < n2 - 1 && k > 0 && k < n3 - 1) {
        const unsigned m = lab[c];
        t1 = t0 + __ldg(bh + m) * (tin[c + 1] + tin[c - 1] + tin[c + s2] + tin[c - s2] + tin[c + s1] + tin[c - s1] - 6.0f * t0)
           + __ldg(perf + m) * (core - t0);
        if (q) t1 += q[c];
        d1 += cem43_step(t0, t1, dt);
        if (slice && j == sel_j) slice[((long long)i * n3 + k) * slice_steps + slice_col] = t1;
        if (points) { const unsigned id = points[c]; if (id) temp_points[(long long)(id - 1) * total_steps + step] = t1; }
    }
    tout[c] = t1;
    dout[c] = d1;
}

struct GrowBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t need(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        cudaError_t e = cudaMalloc(&p, bytes + 256);
        if (e == cudaSuccess) cap = bytes + 256;
        return e;
    }
};
// buffers of the thermal path live per device for the life of the process: a treatment plan calls BHTE dozens of times
// in a row on the same grid (on / off / pause segments of every repetition, CalculateTemperatureEffects.py:347-455)
struct BhteCtx {
    std::mutex mu;
    cudaStream_t st = nullptr;
    GrowBuf t[2], d[2], q, lab, tab, slice, points, tpoints;
};
BhteCtx g_bhte[64];

// CT-derived maps carry up to 6 + 1024 bone materials (BabelIntegrationBASE.py:1241-1244): 16-bit labels on the device
__global__ void labels_to_u16(const uint32_t *__restrict__ in, unsigned short *__restrict__ out, long long n) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) out[c] = (unsigned short)in[c];
}
}  // namespace

extern "C" int bb_bhte_run(int n1, int n2, int n3, int nmat, int nfields, const float *q, const uint32_t *labels, const float *bh,
                           const float *perf, float *temp, float *dose, const int16_t *field_at_step, int64_t total_steps, float dt,
                           float core_temp, int sel_j, int nfactor_monitoring, float *monitor_slice, const uint32_t *monitor_points,
                           int64_t npoints, float *temp_points, int device, double *kernel_ms) {
    BB_REQUIRE(n1 > 2 && n2 > 2 && n3 > 2 && nmat >= 1 && nmat <= 65535 && nfields >= 1, "bad BHTE sizes");
    BB_REQUIRE(q && labels && bh && perf && temp && dose && field_at_step && total_steps >= 0, "null BHTE argument");
    BB_REQUIRE(nfactor_monitoring >= 1, "nFactorMonitoring must be >= 1");
    BB_REQUIRE(!monitor_points || (npoints > 0 && temp_points), "monitoring points need an output table");
    int ndev = bb_device_count();
    if (ndev <= 0) { if (ndev == 0) bb_set_error("no CUDA device (this library has no CPU fallback)"); return BB_ERR_CUDA; }
    BB_REQUIRE(device >= 0 && device < ndev && device < 64, "device %d of %d", device, ndev);
    for (int64_t n = 0; n < total_steps; n++) BB_REQUIRE(field_at_step[n] >= -1 && field_at_step[n] < nfields, "bad field index at step %lld", (long long)n);
    BB_CUDA(cudaSetDevice(device));
    BhteCtx &c = g_bhte[device];
    std::lock_guard<std::mutex> lock(c.mu);
    if (!c.st) BB_CUDA(cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking));
    cudaStream_t st = c.st;
    const size_t cells = (size_t)n1 * n2 * n3;
    const int64_t slice_steps = monitor_slice ? total_steps / nfactor_monitoring : 0;
    for (int b = 0; b < 2; b++) { BB_CUDA(c.t[b].need(cells * 4)); BB_CUDA(c.d[b].need(cells * 4)); }
    BB_CUDA(c.q.need((size_t)nfields * cells * 4));
    BB_CUDA(c.lab.need(cells * 6 + 512));           // uint16 labels behind a uint32 staging copy
    BB_CUDA(c.tab.need((size_t)2 * nmat * 4));
    if (slice_steps) BB_CUDA(c.slice.need((size_t)n1 * n3 * slice_steps * 4));
    if (monitor_points) { BB_CUDA(c.points.need(cells * 4)); BB_CUDA(c.tpoints.need((size_t)npoints * total_steps * 4)); }
    float *dT[2] = { (float *)c.t[0].p, (float *)c.t[1].p }, *dD[2] = { (float *)c.d[0].p, (float *)c.d[1].p };
    unsigned short *dlab = (unsigned short *)c.lab.p;
    uint32_t *dlab32 = (uint32_t *)((char *)c.lab.p + ((cells * 2 + 255) & ~(size_t)255));
    float *dtab = (float *)c.tab.p;
    BB_CUDA(cudaMemcpyAsync(dT[0], temp, cells * 4, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemcpyAsync(dD[0], dose, cells * 4, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemcpyAsync(c.q.p, q, (size_t)nfields * cells * 4, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemcpyAsync(dlab32, labels, cells * 4, cudaMemcpyHostToDevice, st));
    labels_to_u16<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(dlab32, dlab, (long long)cells);
    BB_CUDA(cudaGetLastError());
    BB_CUDA(cudaMemcpyAsync(dtab, bh, (size_t)nmat * 4, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemcpyAsync(dtab + nmat, perf, (size_t)nmat * 4, cudaMemcpyHostToDevice, st));
    if (slice_steps) BB_CUDA(cudaMemsetAsync(c.slice.p, 0, (size_t)n1 * n3 * slice_steps * 4, st));
    if (monitor_points) {
        BB_CUDA(cudaMemcpyAsync(c.points.p, monitor_points, cells * 4, cudaMemcpyHostToDevice, st));
        BB_CUDA(cudaMemsetAsync(c.tpoints.p, 0, (size_t)npoints * total_steps * 4, st));
    }
    cudaEvent_t e0, e1;
    BB_CUDA(cudaEventCreate(&e0));
    BB_CUDA(cudaEventCreate(&e1));
    const dim3 blk(BX, BY, BZ), grid((n3 + BX - 1) / BX, (n2 + BY - 1) / BY, (n1 + BZ - 1) / BZ);
    cudaEventRecord(e0, st);
    int cur = 0;
    for (int64_t n = 0; n < total_steps; n++) {
        const int f = field_at_step[n];
        const bool mon = slice_steps && (n % nfactor_monitoring == 0) && (n / nfactor_monitoring) < slice_steps;
        bhte_kernel<<<grid, blk, 0, st>>>(n1, n2, n3, dT[cur], dT[cur ^ 1], dD[cur], dD[cur ^ 1], dlab, dtab, dtab + nmat,
                                          f >= 0 ? (const float *)c.q.p + (size_t)f * cells : nullptr, core_temp, dt, sel_j,
                                          mon ? (float *)c.slice.p : nullptr, slice_steps, n / nfactor_monitoring,
                                          monitor_points ? (const unsigned *)c.points.p : nullptr, (float *)c.tpoints.p, total_steps, n);
        cur ^= 1;
    }
    cudaEventRecord(e1, st);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess) err = cudaMemcpyAsync(temp, dT[cur], cells * 4, cudaMemcpyDeviceToHost, st);
    if (err == cudaSuccess) err = cudaMemcpyAsync(dose, dD[cur], cells * 4, cudaMemcpyDeviceToHost, st);
    if (err == cudaSuccess && slice_steps) err = cudaMemcpyAsync(monitor_slice, c.slice.p, (size_t)n1 * n3 * slice_steps * 4, cudaMemcpyDeviceToHost, st);
    if (err == cudaSuccess && monitor_points) err = cudaMemcpyAsync(temp_points, c.tpoints.p, (size_t)npoints * total_steps * 4, cudaMemcpyDeviceToHost, st);
    if (err == cudaSuccess) err = cudaStreamSynchronize(st);
    float ms = 0;
    if (err == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (err != cudaSuccess) { bb_set_error("bb_bhte_run: %s", cudaGetErrorString(err)); return BB_ERR_CUDA; }
    if (kernel_ms) *kernel_ms = ms;
    return BB_OK;
}
