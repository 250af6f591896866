// Rayleigh-Sommerfeld forward integral on the GPU; replaces ForwardSimple(cwvnb, center, ds, u0, rf)
// (TranscranialModeling/BabelIntegrationSingle.py:295, BabelIntegrationANNULAR_ARRAY.py:383,411).
//   out[p] = j k/(2 pi) * sum_s ds[s] exp(Im(k) R)/R * u0[s] exp(-j Re(k) R),  R = |rf[p]-center[s]|
// Compute bound (FP32 + MUFU): sources are staged through shared memory as float4 + float2 records
// and every thread accumulates PPT field points so each staged source is reused PPT*blockDim times.
// Per source-point pair: 3 sub, 3 fma (R^2), rsqrt.approx + residual correction of R, sin.approx +
// cos.approx on a phase reduced to one revolution with FMAs (error = rounding of R only), 1 mul
// (amplitude), 6 fma-class accumulations: ~22 FP32 + 3 MUFU issue slots.
// MUFU runs at a quarter of the FP32 rate, so the SFU bounds the kernel at
// 148 SMs x 16 MUFU/clk x 1.9 GHz / 3 = 1.5e12 pairs/s.
#include "common.h"

namespace {
constexpr int RB = 256;      // threads per CTA
constexpr int PPT = 4;       // field points per thread
constexpr int STILE = 512;   // sources per shared-memory tile

template <bool ATT, bool MAXD, bool PERPOINT>
__global__ void __launch_bounds__(RB) rayleigh_kernel(float k_re, float k_im, long long nsrc, const float *__restrict__ center,
                                                       const float *__restrict__ ds, const float2 *__restrict__ u0,
                                                       long long npts, const float *__restrict__ rf, float2 *__restrict__ out,
                                                       float max_distance) {
    __shared__ float4 s_pos[STILE];  // x, y, z, ds
    __shared__ float2 s_u[STILE];
    float px[PPT], py[PPT], pz[PPT], ar[PPT], ai[PPT];
    const double krd = (double)k_re * 0.15915494309189535;       // wavenumber in revolutions per metre, hi + lo
    const float kr_hi = (float)krd, kr_lo = (float)(krd - (double)kr_hi);
    long long pidx[PPT];
#pragma unroll
    for (int t = 0; t < PPT; t++) {
        pidx[t] = ((long long)blockIdx.x * PPT + t) * RB + threadIdx.x;
        const long long pp = pidx[t] < npts ? pidx[t] : npts - 1;
        px[t] = rf[3 * pp]; py[t] = rf[3 * pp + 1]; pz[t] = rf[3 * pp + 2];
        ar[t] = 0.f; ai[t] = 0.f;
    }
    for (long long s0 = 0; s0 < nsrc; s0 += STILE) {
        const int ns = (int)min((long long)STILE, nsrc - s0);
        __syncthreads();
        for (int s = threadIdx.x; s < ns; s += RB) {
            const long long g = s0 + s;
            s_pos[s] = make_float4(center[3 * g], center[3 * g + 1], center[3 * g + 2], ds[g]);
            if (!PERPOINT) s_u[s] = u0[g];
        }
        __syncthreads();
#pragma unroll 4
        for (int s = 0; s < ns; s++) {
            const float4 c = s_pos[s];
            float2 u;
            if (!PERPOINT) u = s_u[s];
#pragma unroll
            for (int t = 0; t < PPT; t++) {
                const float dx = c.x - px[t], dy = c.y - py[t], dz = c.z - pz[t];
                const float r2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                float rinv;
                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rinv) : "f"(r2));   // R = 0 -> inf, like the reference's 1/R
                // R = sqrt(r2) to ~0.5 ulp: the phase R k is 1e2..1e3 rad, so an ulp of R is what sets the error of the sum.
                // (the amplitude keeps the 2-ulp rinv: 1e-7 relative)
                const float R0 = r2 * rinv;
                const float R = fmaf(fmaf(-R0, R0, r2), 0.5f * rinv, R0);
                if (MAXD && R > max_distance) continue;
                if (PERPOINT) u = u0[(pidx[t] < npts ? pidx[t] : npts - 1) * nsrc + s0 + s];
                float amp = c.w * rinv;
                if (ATT) amp *= __expf(R * k_im);
                // phase in revolutions, reduced with fused multiply-adds against a two-term k/(2 pi): the only
                // error left is the rounding of R itself
                const float n = rintf(R * kr_hi);
                const float fr = fmaf(R, kr_hi, -n) + R * kr_lo;
                const float sn = __sinf(fr * 6.283185307f), cs = __cosf(fr * 6.283185307f);
                ar[t] += amp * (u.x * cs + u.y * sn);
                ai[t] += amp * (u.y * cs - u.x * sn);
            }
        }
    }
    const float inv2pi = 0.15915494309189535f;
#pragma unroll
    for (int t = 0; t < PPT; t++) {
        if (pidx[t] < npts)
            out[pidx[t]] = make_float2((-ar[t] * k_im - ai[t] * k_re) * inv2pi, (ar[t] * k_re - ai[t] * k_im) * inv2pi);
    }
}
}  // namespace

// Device buffers, stream and events of the Rayleigh path live per device for the life of the process: the transducer
// models call ForwardSimple many times in a row (phase programming point by point, then the whole grid;
// BabelIntegrationANNULAR_ARRAY.py:376-420), and nothing is allocated or freed per call unless a buffer has to grow.
#include <mutex>
namespace {
struct GrowBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t need(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        const size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
};
struct RayleighCtx {
    std::mutex mu;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    GrowBuf center, ds, u0, rf, out;
};
RayleighCtx g_rl[64];
}  // namespace

extern "C" int bb_rayleigh_forward(float k_re, float k_im, int64_t nsrc, const float *center, const float *ds,
                                   const float *u0_reim, int64_t npts, const float *rf, float *out_reim,
                                   float max_distance, int64_t u0_step, int device, double *kernel_ms) {
    BB_REQUIRE(nsrc > 0 && npts > 0 && center && ds && u0_reim && rf && out_reim, "bad Rayleigh arguments");
    BB_REQUIRE(u0_step == 0 || u0_step == nsrc, "u0step must equal the number of sources");
    int ndev = bb_device_count();
    if (ndev <= 0) { if (ndev == 0) bb_set_error("no CUDA device (this library has no CPU fallback)"); return BB_ERR_CUDA; }
    BB_REQUIRE(device >= 0 && device < ndev && device < 64, "device %d of %d", device, ndev);
    BB_CUDA(cudaSetDevice(device));
    RayleighCtx &c = g_rl[device];
    std::lock_guard<std::mutex> lock(c.mu);
    if (!c.st) {
        BB_CUDA(cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking));
        BB_CUDA(cudaEventCreate(&c.e0));
        BB_CUDA(cudaEventCreate(&c.e1));
    }
    cudaStream_t st = c.st;
    const size_t nu = u0_step ? (size_t)npts * nsrc : (size_t)nsrc;
    BB_CUDA(c.center.need((size_t)nsrc * 12));
    BB_CUDA(c.ds.need((size_t)nsrc * 4));
    BB_CUDA(c.u0.need(nu * 8));
    BB_CUDA(c.rf.need((size_t)npts * 12));
    BB_CUDA(c.out.need((size_t)npts * 8));
    float *d_center = (float *)c.center.p, *d_ds = (float *)c.ds.p, *d_rf = (float *)c.rf.p;
    float2 *d_u0 = (float2 *)c.u0.p, *d_out = (float2 *)c.out.p;
    BB_CUDA(cudaMemcpyAsync(d_center, center, (size_t)nsrc * 12, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemcpyAsync(d_ds, ds, (size_t)nsrc * 4, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemcpyAsync(d_u0, u0_reim, nu * 8, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemcpyAsync(d_rf, rf, (size_t)npts * 12, cudaMemcpyHostToDevice, st));
    const unsigned grid = (unsigned)((npts + (long long)RB * PPT - 1) / ((long long)RB * PPT));
    const bool att = k_im != 0.f, maxd = max_distance > 0.f, pp = u0_step != 0;
    BB_CUDA(cudaEventRecord(c.e0, st));
#define BB_RL(A, M, Q) rayleigh_kernel<A, M, Q><<<grid, RB, 0, st>>>(k_re, k_im, nsrc, d_center, d_ds, d_u0, npts, d_rf, d_out, max_distance)
    if (pp) { if (att) { if (maxd) BB_RL(true, true, true); else BB_RL(true, false, true); } else { if (maxd) BB_RL(false, true, true); else BB_RL(false, false, true); } }
    else { if (att) { if (maxd) BB_RL(true, true, false); else BB_RL(true, false, false); } else { if (maxd) BB_RL(false, true, false); else BB_RL(false, false, false); } }
#undef BB_RL
    BB_CUDA(cudaGetLastError());
    BB_CUDA(cudaEventRecord(c.e1, st));
    BB_CUDA(cudaMemcpyAsync(out_reim, d_out, (size_t)npts * 8, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    BB_CUDA(cudaEventElapsedTime(&ms, c.e0, c.e1));
    if (kernel_ms) *kernel_ms = ms;
    return BB_OK;
}
