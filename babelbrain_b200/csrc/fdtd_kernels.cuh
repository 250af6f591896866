// Auxiliary kernels of the FDTD path: source injection, sensor sampling, format conversion.
#pragma once
#include "fdtd_cell.cuh"

// ------------------------------------------------------------------------------------------
// sources, sensors, bookkeeping
// ------------------------------------------------------------------------------------------
// one thread per source cell; sf is [nt_src][nsrc] (row n contiguous).  A source in one of the slab's boundary
// planes also updates the copy the slab neighbour holds in its halo (NVLink halo push, DevParams::peerV).
__device__ __forceinline__ void push_source_cell(const DevParams &p, float *const *peer, int ncomp, const int *comp, const float *val, long long q) {
    const long long ipl = q / p.plane;
    const int i = (int)ipl - 2 + p.i0;
    const long long col = q - ipl * p.plane;
#pragma unroll
    for (int side = 0; side < 2; side++) {
        if (!peer[side]) continue;
        const bool mine = side == 0 ? (i < p.i0 + 2) : (i >= p.i1 - 2);
        if (!mine) continue;
        const long long qn = (long long)(p.peer_plane[side] + (side == 0 ? i - p.i0 : i - (p.i1 - 2))) * p.plane + col;
        for (int c = 0; c < ncomp; c++) peer[side][comp[c] * p.peer_vol[side] + qn] = val[c];
    }
}

__global__ void source_kernel(const DevParams p, int type_source, long long ncells, const long long *__restrict__ cell,
                              const int *__restrict__ row, const float *__restrict__ ox,
                              const float *__restrict__ oy, const float *__restrict__ oz,
                              const float *__restrict__ sf_row, const float *__restrict__ tone_ac,
                              const float *__restrict__ tone_as, float env_sin, float env_cos) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= ncells) return;
    // table row of this time step, or the continuous-wave source evaluated in place:
    // ramp(n) A sin(w t_n + phi) = [ramp sin(w t_n)] A cos(phi) + [ramp cos(w t_n)] A sin(phi)
    const float v = tone_ac ? fmaf(env_sin, tone_ac[row[s]], env_cos * tone_as[row[s]]) : sf_row[row[s]];
    const long long q = cell[s];
    if (type_source >= 2) {
        const float w = v * ox[s];
        if (type_source == 2) { p.S[0][q] += w; p.S[1][q] += w; p.S[2][q] += w; }
        else { p.S[0][q] = w; p.S[1][q] = w; p.S[2][q] = w; }
        if (p.peerS[0] || p.peerS[1]) {
            const int comp[1] = { 0 };
            const float val[1] = { p.S[0][q] };
            push_source_cell(p, p.peerS, 1, comp, val, q);
            __threadfence_system();
        }
    } else {
        if (type_source == 0) { p.V[0][q] += v * ox[s]; p.V[1][q] += v * oy[s]; p.V[2][q] += v * oz[s]; }
        else { p.V[0][q] = v * ox[s]; p.V[1][q] = v * oy[s]; p.V[2][q] = v * oz[s]; }
        if (p.peerV[0] || p.peerV[1]) {
            const int comp[3] = { 0, 1, 2 };
            const float val[3] = { p.V[0][q], p.V[1][q], p.V[2][q] };
            push_source_cell(p, p.peerV, 3, comp, val, q);
            __threadfence_system();
        }
    }
}

template <typename LT>
__device__ __forceinline__ float field_value(const DevParams &p, int map, long long q) {
    switch (map) {
    case BB_MAP_ALLV: {
        const float a = p.V[0][q], b = p.V[1][q], c = p.V[2][q];
        return sqrtf(a * a + b * b + c * c);
    }
    case BB_MAP_VX: case BB_MAP_VY: case BB_MAP_VZ: return p.V[map - BB_MAP_VX][q];
    case BB_MAP_PRESSURE: {
        const unsigned m = reinterpret_cast<const LT *>(p.lab)[q] & LabelTraits<LT>::MASK;
        return -__ldg(&p.coef[m].K) * p.Pr[q];
    }
    default: return p.S[map - BB_MAP_SXX][q];
    }
}

// out layout on the device: [selected map][sample][sensor]
template <typename LT>
__global__ void sensor_kernel(const DevParams p, unsigned sel_maps, long long nsensors, const long long *__restrict__ cell,
                              float *__restrict__ out, long long nsamples, long long sample) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsensors) return;
    const long long q = cell[s];
    int slot = 0;
    for (int m = 0; m < BB_MAP_COUNT; m++) {
        if (!(sel_maps & (1u << m))) continue;
        out[((long long)slot * nsamples + sample) * nsensors + s] = field_value<LT>(p, m, q);
        slot++;
    }
}

// (sample, sensor) -> (sensor, sample) for the sensors [s0, s0 + ns): a CTA stages 256 sensors x nsamples through shared
// memory, so both the reads (consecutive sensors of one sample) and the writes (the 256 x nsamples floats of the CTA's
// sensors are one contiguous run of the output) are coalesced.  out points at the row of sensor s0.
constexpr int BB_ST_SENSORS = 256;
__global__ void __launch_bounds__(BB_ST_SENSORS) sensor_transpose_kernel(const float *__restrict__ in, float *__restrict__ out, long long nsensors,
                                                                          int nsamples, long long s0, long long ns) {
    extern __shared__ float st_tile[];              // [nsamples][BB_ST_SENSORS + 1]
    const long long base = (long long)blockIdx.x * BB_ST_SENSORS;
    const int nhere = (int)min((long long)BB_ST_SENSORS, ns - base);
    for (int t = 0; t < nsamples; t++)
        if ((int)threadIdx.x < nhere) st_tile[t * (BB_ST_SENSORS + 1) + threadIdx.x] = in[(long long)t * nsensors + s0 + base + threadIdx.x];
    __syncthreads();
    float *o = out + base * nsamples;
    for (int e = threadIdx.x; e < nhere * nsamples; e += BB_ST_SENSORS) {
        const int sl = e / nsamples, t = e - sl * nsamples;
        o[e] = st_tile[t * (BB_ST_SENSORS + 1) + sl];
    }
}

// The same for one slab of a multi-GPU run, written straight into the whole-grid table in page-locked host memory: the
// slab's rows form runs of consecutive rows (one per (j,k) line, slab.merge_sensor_runs); run u holds the slab's rows
// [src_row[u], src_row[u+1]) and starts at row dst_row[u] of the table.  Rows of one run stay contiguous, so the PCIe
// writes are as coalesced as the device-memory ones above.
__global__ void __launch_bounds__(BB_ST_SENSORS) sensor_transpose_scatter_kernel(const float *__restrict__ in, float *__restrict__ table, long long nsensors,
                                                                                  int nsamples, const long long *__restrict__ src_row,
                                                                                  const long long *__restrict__ dst_row, long long nruns) {
    extern __shared__ float st_tile[];              // [nsamples][BB_ST_SENSORS + 1]
    __shared__ long long grow[BB_ST_SENSORS];
    const long long base = (long long)blockIdx.x * BB_ST_SENSORS;
    const int nhere = (int)min((long long)BB_ST_SENSORS, nsensors - base);
    if ((int)threadIdx.x < nhere) {
        const long long s = base + threadIdx.x;
        long long lo = 0, hi = nruns;               // last run with src_row <= s
        while (hi - lo > 1) { const long long mid = (lo + hi) >> 1; if (src_row[mid] <= s) lo = mid; else hi = mid; }
        grow[threadIdx.x] = dst_row[lo] + (s - src_row[lo]);
    }
    for (int t = 0; t < nsamples; t++)
        if ((int)threadIdx.x < nhere) st_tile[t * (BB_ST_SENSORS + 1) + threadIdx.x] = in[(long long)t * nsensors + base + threadIdx.x];
    __syncthreads();
    for (int e = threadIdx.x; e < nhere * nsamples; e += BB_ST_SENSORS) {
        const int sl = e / nsamples, t = e - sl * nsamples;
        table[grow[sl] * nsamples + t] = st_tile[t * (BB_ST_SENSORS + 1) + sl];
    }
}

// Single-bin DFT of each sensor's trace (in: [sample][sensor]), its angle and the largest sample, scattered to the dense
// (nown, n2, n3) volumes of the slab.  tw = nsamples (cos, -sin) pairs of the bin, rounded from double on the host.
__global__ void phase_data_kernel(const DevParams p, const float *__restrict__ in, const long long *__restrict__ cell, long long nsensors,
                                  int nsamples, const float2 *__restrict__ tw, float scale, float2 *__restrict__ fourier,
                                  float *__restrict__ phase, float *__restrict__ peak) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsensors) return;
    float re = 0.f, im = 0.f, mx = -INFINITY;
    for (int n = 0; n < nsamples; n++) {
        const float x = in[(long long)n * nsensors + s];
        const float2 w = __ldg(tw + n);
        re = fmaf(x, w.x, re);
        im = fmaf(x, w.y, im);
        mx = fmaxf(mx, x);
    }
    const long long q = cell[s];                       // pitched local index (ipl * n2 + j) * pitch + k, ipl = i - i0 + 2
    const long long ipl = q / p.plane, rem = q - ipl * p.plane;
    const long long j = rem / p.pitch, k = rem - j * p.pitch;
    const long long o = ((ipl - 2) * p.n2 + j) * p.n3 + k;
    fourier[o] = make_float2(re * scale, im * scale);
    if (phase) phase[o] = atan2f(im, re);
    if (peak) peak[o] = mx;
}

// uint32 host labels (dense rows of n3) -> LT labels (pitched rows), reflector -> top bit
template <typename LT>
__global__ void label_convert_kernel(const uint32_t *__restrict__ in, const uint32_t *__restrict__ refl, LT *__restrict__ out,
                                     long long nrows, int n3, int pitch, int nmat, int *__restrict__ bad) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long r = t / pitch;
    const int k = (int)(t - r * pitch);
    if (r >= nrows) return;
    unsigned v = 0;
    if (k < n3) {
        v = in[r * n3 + k];
        if (v >= (unsigned)nmat) { atomicExch(bad, 1); v = 0; }
        if (refl && refl[r * n3 + k]) v |= LabelTraits<LT>::REFL;
    }
    out[r * pitch + k] = (LT)v;
}

// `nsrc` rows of SourceFunctions (nsrc, nt) with row stride -> float32 columns of out[nt][out_stride]; 32x32 smem tile transpose
template <typename T>
__global__ void srcfun_transpose_kernel(const T *__restrict__ in, long long row_stride, float *__restrict__ out, int nsrc, int nt,
                                        int out_stride) {
    __shared__ float tile[32][33];
    const int t0 = blockIdx.x * 32, s0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int s = s0 + r, t = t0 + threadIdx.x;
        tile[r][threadIdx.x] = (s < nsrc && t < nt) ? (float)in[(long long)s * row_stride + t] : 0.0f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int t = t0 + r, s = s0 + threadIdx.x;
        if (t < nt && s < nsrc) out[(long long)t * out_stride + s] = tile[threadIdx.x][r];
    }
}

// dense (nown, n2, n3) result from a pitched owned-planes array; mode 0 copy, 1 sqrt(x/nacc), 2 sqrt(x)
__global__ void finalize_map_kernel(const float *__restrict__ in, float *__restrict__ out, long long nrows, int n3, int pitch,
                                    int mode, float inv_nacc) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long r = t / n3;
    const int k = (int)(t - r * n3);
    if (r >= nrows) return;
    float v = in[r * pitch + k];
    if (mode == 1) v = sqrtf(v * inv_nacc);
    else if (mode == 2) v = sqrtf(v);
    out[t] = v;
}

// rows [r0, r0 + nrows) of the owned planes
template <typename LT>
__global__ void finalize_pressure_kernel(const DevParams p, float *__restrict__ out, long long r0, long long nrows) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long r = t / p.n3;
    const int k = (int)(t - r * p.n3);
    if (r >= nrows) return;
    const long long q = 2 * p.plane + (r0 + r) * p.pitch + k;
    out[t] = field_value<LT>(p, BB_MAP_PRESSURE, q);
}
