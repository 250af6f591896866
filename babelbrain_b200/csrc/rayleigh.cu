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
// rayleigh_kernel2 below is the production kernel (packed FP32 pairs, 16 issue slots per pair, 85 % of that bound);
// this scalar kernel serves the per-point-u0 form and BB_RAYLEIGH_SCALAR=1 comparison runs.
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

// ---- the same sum on packed FP32 pairs (sm_100 FFMA2 / FMUL2 / FADD2: one issue slot for two field points)
// The scalar kernel above issues ~26 instructions per source-point pair and is bound by instruction issue (4 per clock and
// SM) as much as by the special-function unit; the conversion behind rintf() runs on that unit as well, which made four
// of its operations per pair.  Here every thread carries its field points as pairs in 64-bit registers, the sources sit
// in shared memory already duplicated into both halves ((x,x) (y,y) (z,z) and ds*u0 as (ux,ux) (uy,uy) (-ux,-ux)), the
// revolution count is rounded with the 1.5*2^23 constant on the FMA pipe, and only rsqrt / sin / cos (and ex2 with
// attenuation) remain scalar: ~15 issue slots and 3 special-function operations per pair.
typedef unsigned long long p2;
__device__ __forceinline__ p2 pk(float a, float b) { p2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up(p2 v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ p2 ffma2(p2 a, p2 b, p2 c) { p2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ p2 fmul2(p2 a, p2 b) { p2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ p2 fsub2(p2 a, p2 b) { p2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ p2 neg2(p2 a) { float x, y; up(a, x, y); return pk(-x, -y); }       // folds into the consumer's operand modifier
__device__ __forceinline__ p2 dup(float a) { return pk(a, a); }

template <bool ATT, bool MAXD, int NP2>      // NP2 pairs of field points per thread
__global__ void __launch_bounds__(RB) rayleigh_kernel2(float k_re, float k_im, long long nsrc, const float *__restrict__ center,
                                                        const float *__restrict__ ds, const float2 *__restrict__ u0,
                                                        long long npts, const float *__restrict__ rf, float2 *__restrict__ out,
                                                        float max_distance) {
    __shared__ ulonglong2 s_a[STILE];   // (x,x) (y,y)
    __shared__ ulonglong2 s_b[STILE];   // (z,z) (ux,ux)     u = ds * u0
    __shared__ ulonglong2 s_c[STILE];   // (uy,uy) (-ux,-ux)
    constexpr int NPT = 2 * NP2;
    p2 px[NP2], py[NP2], pz[NP2], ar[NP2], ai[NP2];
    const double krd = (double)k_re * 0.15915494309189535;       // wavenumber in revolutions per metre, hi + lo
    const float kr_hi = (float)krd, kr_lo = (float)(krd - (double)kr_hi);
    const p2 KRHI = dup(kr_hi), KRLO = dup(kr_lo), MAGIC = dup(12582912.0f), HALF = dup(0.5f), TWOPI = dup(6.283185307f);
    const p2 KIM = dup(k_im * 1.4426950408889634f);              // exp(R k_im) = 2^(R k_im log2 e)
    long long pidx[NPT];
#pragma unroll
    for (int q = 0; q < NP2; q++) {
        float x[2], y[2], z[2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int t = 2 * q + e;
            pidx[t] = ((long long)blockIdx.x * NPT + t) * RB + threadIdx.x;
            const long long pp = pidx[t] < npts ? pidx[t] : npts - 1;
            x[e] = rf[3 * pp]; y[e] = rf[3 * pp + 1]; z[e] = rf[3 * pp + 2];
        }
        px[q] = pk(x[0], x[1]); py[q] = pk(y[0], y[1]); pz[q] = pk(z[0], z[1]);
        ar[q] = 0ull; ai[q] = 0ull;
    }
    for (long long s0 = 0; s0 < nsrc; s0 += STILE) {
        const int ns = (int)min((long long)STILE, nsrc - s0);
        __syncthreads();
        for (int s = threadIdx.x; s < ns; s += RB) {
            const long long g = s0 + s;
            const float w = ds[g];
            const float2 u = u0[g];
            const float ux = w * u.x, uy = w * u.y;
            s_a[s] = make_ulonglong2(dup(center[3 * g]), dup(center[3 * g + 1]));
            s_b[s] = make_ulonglong2(dup(center[3 * g + 2]), dup(ux));
            s_c[s] = make_ulonglong2(dup(uy), dup(-ux));
        }
        __syncthreads();
#pragma unroll 2
        for (int s = 0; s < ns; s++) {
            const ulonglong2 a = s_a[s], b = s_b[s], c = s_c[s];
#pragma unroll
            for (int q = 0; q < NP2; q++) {
                const p2 dx = fsub2(a.x, px[q]), dy = fsub2(a.y, py[q]), dz = fsub2(b.x, pz[q]);
                const p2 r2 = ffma2(dx, dx, ffma2(dy, dy, fmul2(dz, dz)));
                float r2a, r2b, ia, ib;
                up(r2, r2a, r2b);
                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(ia) : "f"(r2a));   // R = 0 -> inf, like the reference's 1/R
                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(ib) : "f"(r2b));
                const p2 rinv = pk(ia, ib);
                // R = sqrt(r2) to ~0.5 ulp (one Newton step on the 2-ulp rsqrt): the phase R k is 1e2..1e3 rad, so an
                // ulp of R is what sets the error of the sum; the amplitude keeps the 2-ulp rinv
                const p2 R0 = fmul2(r2, rinv);
                const p2 R = ffma2(ffma2(neg2(R0), R0, r2), fmul2(rinv, HALF), R0);
                p2 amp = rinv;
                if (MAXD) {
                    float Ra, Rb;
                    up(R, Ra, Rb);
                    amp = pk(Ra > max_distance ? 0.f : ia, Rb > max_distance ? 0.f : ib);
                }
                if (ATT) {
                    float ea, eb;
                    up(fmul2(R, KIM), ea, eb);
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ea) : "f"(ea));
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(eb) : "f"(eb));
                    amp = fmul2(amp, pk(ea, eb));
                }
                // phase in revolutions, reduced with fused multiply-adds against a two-term k/(2 pi); the revolution count is
                // rounded by adding 1.5 * 2^23 (exact below 2^22 revolutions), not by a conversion instruction
                const p2 t = ffma2(R, KRHI, MAGIC);
                const p2 nn = fsub2(MAGIC, t);                                // -(whole revolutions)
                const p2 fr = ffma2(R, KRLO, ffma2(R, KRHI, nn));
                float xa, xb;
                up(fmul2(fr, TWOPI), xa, xb);
                const p2 sn = pk(__sinf(xa), __sinf(xb)), cs = pk(__cosf(xa), __cosf(xb));
                const p2 c2 = fmul2(cs, amp), s2 = fmul2(sn, amp);
                ar[q] = ffma2(c.x, s2, ffma2(b.y, c2, ar[q]));                // += amp (ux cs + uy sn)
                ai[q] = ffma2(c.y, s2, ffma2(c.x, c2, ai[q]));                // += amp (uy cs - ux sn)
            }
        }
    }
    const float inv2pi = 0.15915494309189535f;
#pragma unroll
    for (int q = 0; q < NP2; q++) {
        float r[2], i[2];
        up(ar[q], r[0], r[1]); up(ai[q], i[0], i[1]);
#pragma unroll
        for (int e = 0; e < 2; e++)
            if (pidx[2 * q + e] < npts)
                out[pidx[2 * q + e]] = make_float2((-r[e] * k_im - i[e] * k_re) * inv2pi, (r[e] * k_re - i[e] * k_im) * inv2pi);
    }
}
#ifndef BB_RL_NP2
#define BB_RL_NP2 2
#endif
constexpr int RL_NP2 = BB_RL_NP2;      // pairs of field points per thread in rayleigh_kernel2
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
    const unsigned grid2 = (unsigned)((npts + (long long)RB * 2 * RL_NP2 - 1) / ((long long)RB * 2 * RL_NP2));
    const bool att = k_im != 0.f, maxd = max_distance > 0.f, pp = u0_step != 0;
    static const bool scalar_kernel = getenv("BB_RAYLEIGH_SCALAR") != nullptr;      // comparison runs only
    BB_CUDA(cudaEventRecord(c.e0, st));
#define BB_RL2(A, M) rayleigh_kernel2<A, M, RL_NP2><<<grid2, RB, 0, st>>>(k_re, k_im, nsrc, d_center, d_ds, d_u0, npts, d_rf, d_out, max_distance)
    if (!pp && !scalar_kernel) { if (att) { if (maxd) BB_RL2(true, true); else BB_RL2(true, false); } else { if (maxd) BB_RL2(false, true); else BB_RL2(false, false); } }
    else
#undef BB_RL2
#define BB_RL(A, M, Q) rayleigh_kernel<A, M, Q><<<grid, RB, 0, st>>>(k_re, k_im, nsrc, d_center, d_ds, d_u0, npts, d_rf, d_out, max_distance)
    {
    if (pp) { if (att) { if (maxd) BB_RL(true, true, true); else BB_RL(true, false, true); } else { if (maxd) BB_RL(false, true, true); else BB_RL(false, false, true); } }
    else { if (att) { if (maxd) BB_RL(true, true, false); else BB_RL(true, false, false); } else { if (maxd) BB_RL(false, true, false); else BB_RL(false, false, false); } }
    }
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
