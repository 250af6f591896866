// kernel_variant = 1 ("direct"): one thread per cell, operands straight from global memory through
// L1/L2, no shared memory, no tile flags.  It is the simple statement of the two half-steps on the
// device data layout: the ablation baseline for the TMA-pipelined kernels and
// the bisecting aid when they disagree with the oracle.
#pragma once
#include "fdtd_cell.cuh"

namespace direct {

__device__ __forceinline__ float dbwd(const float *__restrict__ f, long long q, long long st, const AxisCoef &c) {
    return D4C(c.cab, c.cbb, f[q], f[q - st], f[q + st], f[q - 2 * st]);
}
__device__ __forceinline__ float dfwd(const float *__restrict__ f, long long q, long long st, const AxisCoef &c) {
    return D4C(c.caf, c.cbf, f[q + st], f[q], f[q + 2 * st], f[q - st]);
}

// The arrays carry two halo planes in i, and rows/columns are addressed with j-1.. j+2 / k-1..k+2:
// the first/last two rows and columns would run outside the allocation, so differences whose
// coefficients are zero are still formed from in-bounds (clamped) addresses.
template <typename LT, bool ACC>
__global__ void __launch_bounds__(256) stress_direct(const DevParams p, int ib) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int i = ib + blockIdx.z;
    if (k >= p.n3 || j >= p.n2) return;
    const bool pml = in_pml1(i, p.n1, p.P) || in_pml1(j, p.n2, p.P) || in_pml1(k, p.n3, p.P);
    if (pml && !(i < p.n1 - 1 && j < p.n2 - 1 && k < p.n3 - 1)) return;
    const long long s1 = p.plane, s2 = p.pitch;
    const long long q = ((long long)(i - p.i0 + 2) * p.n2 + j) * s2 + k;
    const LT *__restrict__ lab = reinterpret_cast<const LT *>(p.lab);
    const unsigned MSK = LabelTraits<LT>::MASK;
    const unsigned l0 = lab[q];
    const bool refl = (l0 & LabelTraits<LT>::REFL) != 0;
    const MatCoef c = load_coef_global(p.coef, l0 & MSK);
    const AxisCoef ci = load_axis(p.axI, i), cj = load_axis(p.axJ, j), ck = load_axis(p.axK, k);
    const float *__restrict__ Vx = p.V[0], *__restrict__ Vy = p.V[1], *__restrict__ Vz = p.V[2];
    // interior cells are at least P >= 2 away from every face; PML cells next to a face have zero
    // coefficients on the taps that would leave the volume, but the addresses must stay legal
    const long long j2m = j >= 2 ? 2 * s2 : (j >= 1 ? s2 : 0), j1m = j >= 1 ? s2 : 0;
    const long long j1p = j + 1 < p.n2 ? s2 : 0, j2p = j + 2 < p.n2 ? 2 * s2 : j1p;
    const long long k2m = k >= 2 ? 2 : (k >= 1 ? 1 : 0), k1m = k >= 1 ? 1 : 0;
    const long long k1p = k + 1 < p.n3 ? 1 : 0, k2p = k + 2 < p.n3 ? 2 : k1p;
    float D[9];
    D[0] = D4C(ci.cab, ci.cbb, Vx[q], Vx[q - s1], Vx[q + s1], Vx[q - 2 * s1]);
    D[1] = D4C(cj.cab, cj.cbb, Vy[q], Vy[q - j1m], Vy[q + j1p], Vy[q - j2m]);
    D[2] = D4C(ck.cab, ck.cbb, Vz[q], Vz[q - k1m], Vz[q + k1p], Vz[q - k2m]);
    D[3] = D4C(ci.caf, ci.cbf, Vy[q + s1], Vy[q], Vy[q + 2 * s1], Vy[q - s1]);
    D[4] = D4C(cj.caf, cj.cbf, Vx[q + j1p], Vx[q], Vx[q + j2p], Vx[q - j1m]);
    D[5] = D4C(ci.caf, ci.cbf, Vz[q + s1], Vz[q], Vz[q + 2 * s1], Vz[q - s1]);
    D[6] = D4C(ck.caf, ck.cbf, Vx[q + k1p], Vx[q], Vx[q + k2p], Vx[q - k1m]);
    D[7] = D4C(cj.caf, cj.cbf, Vz[q + j1p], Vz[q], Vz[q + j2p], Vz[q - j1m]);
    D[8] = D4C(ck.caf, ck.cbf, Vy[q + k1p], Vy[q], Vy[q + k2p], Vy[q - k1m]);
    // edge rigidities / relaxation (neighbour labels at +1 in each direction)
    const unsigned mi = lab[q + s1] & MSK, mj = lab[q + j1p] & MSK, mk = lab[q + k1p] & MSK;
    const unsigned mij = lab[q + s1 + j1p] & MSK, mik = lab[q + s1 + k1p] & MSK, mjk = lab[q + j1p + k1p] & MSK;
    const float igi = __ldg(&p.coef[mi].invG), igj = __ldg(&p.coef[mj].invG), igk = __ldg(&p.coef[mk].invG);
    const float rigxy = rigidity4(c.invG, igi, igj, __ldg(&p.coef[mij].invG));
    const float rigxz = rigidity4(c.invG, igi, igk, __ldg(&p.coef[mik].invG));
    const float rigyz = rigidity4(c.invG, igj, igk, __ldg(&p.coef[mjk].invG));
    float s[6];
#pragma unroll
    for (int n = 0; n < 6; n++) s[n] = p.S[n][q];
    if (pml) {
        const PmlCell pc = make_pml_cell(p, i, j, k);
        stress_pml<false>(p, pc, c.M, c.L, rigxy, rigxz, rigyz, D, s);
        if (refl) { s[0] = s[1] = s[2] = s[3] = s[4] = s[5] = 0.0f; }
#pragma unroll
        for (int n = 0; n < 6; n++) p.S[n][q] = s[n];
        return;
    }
    const bool att = attenuates(c);
    float r0 = 0.f, r1 = 0.f, r2 = 0.f, pr = p.Pr[q];
    if (att) { r0 = p.R[0][q]; r1 = p.R[1][q]; r2 = p.R[2][q]; }
    stress_normal_interior(c, p.dt, att, D[0], D[1], D[2], s[0], s[1], s[2], r0, r1, r2, pr);
    if (att) { p.R[0][q] = r0; p.R[1][q] = r1; p.R[2][q] = r2; }
    const float ti = __ldg(&p.coef[mi].tauS), tj = __ldg(&p.coef[mj].tauS), tk = __ldg(&p.coef[mk].tauS);
    if (rigxy != 0.0f) {
        const float te = 0.25f * (c.tauS + ti + tj + __ldg(&p.coef[mij].tauS));
        float r = te != 0.0f ? p.R[3][q] : 0.f;
        stress_shear_interior(c, p.dt, rigxy, te, D[3] + D[4], s[3], r);
        if (te != 0.0f) p.R[3][q] = r;
    }
    if (rigxz != 0.0f) {
        const float te = 0.25f * (c.tauS + ti + tk + __ldg(&p.coef[mik].tauS));
        float r = te != 0.0f ? p.R[4][q] : 0.f;
        stress_shear_interior(c, p.dt, rigxz, te, D[5] + D[6], s[4], r);
        if (te != 0.0f) p.R[4][q] = r;
    }
    if (rigyz != 0.0f) {
        const float te = 0.25f * (c.tauS + tj + tk + __ldg(&p.coef[mjk].tauS));
        float r = te != 0.0f ? p.R[5][q] : 0.f;
        stress_shear_interior(c, p.dt, rigyz, te, D[7] + D[8], s[5], r);
        if (te != 0.0f) p.R[5][q] = r;
    }
    if (refl) { s[0] = s[1] = s[2] = s[3] = s[4] = s[5] = 0.0f; pr = 0.0f; }
#pragma unroll
    for (int n = 0; n < 6; n++) p.S[n][q] = s[n];
    p.Pr[q] = pr;
    if (ACC) {
        const long long qa = q - 2 * s1;
#pragma unroll
        for (int n = 0; n < 6; n++) accumulate(p, BB_MAP_SXX + n, qa, s[n], false);
        accumulate(p, BB_MAP_PRESSURE, qa, -c.K * pr, false);
    }
}

template <typename LT, bool ACC>
__global__ void __launch_bounds__(256) particle_direct(const DevParams p, int ib) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int i = ib + blockIdx.z;
    if (k >= p.n3 || j >= p.n2) return;
    const bool pml = in_pml1(i, p.n1, p.P) || in_pml1(j, p.n2, p.P) || in_pml1(k, p.n3, p.P);
    if (pml && !(i < p.n1 - 1 && j < p.n2 - 1 && k < p.n3 - 1)) return;
    const long long s1 = p.plane, s2 = p.pitch;
    const long long q = ((long long)(i - p.i0 + 2) * p.n2 + j) * s2 + k;
    const LT *__restrict__ lab = reinterpret_cast<const LT *>(p.lab);
    const unsigned MSK = LabelTraits<LT>::MASK;
    const unsigned l0 = lab[q];
    const AxisCoef ci = load_axis(p.axI, i), cj = load_axis(p.axJ, j), ck = load_axis(p.axK, k);
    const long long j2m = j >= 2 ? 2 * s2 : (j >= 1 ? s2 : 0), j1m = j >= 1 ? s2 : 0;
    const long long j1p = j + 1 < p.n2 ? s2 : 0, j2p = j + 2 < p.n2 ? 2 * s2 : j1p;
    const long long k2m = k >= 2 ? 2 : (k >= 1 ? 1 : 0), k1m = k >= 1 ? 1 : 0;
    const long long k1p = k + 1 < p.n3 ? 1 : 0, k2p = k + 2 < p.n3 ? 2 : k1p;
    const float b0 = __ldg(&p.coef[l0 & MSK].B);
    const float bx = 0.5f * (b0 + __ldg(&p.coef[lab[q + s1] & MSK].B));
    const float by = 0.5f * (b0 + __ldg(&p.coef[lab[q + j1p] & MSK].B));
    const float bz = 0.5f * (b0 + __ldg(&p.coef[lab[q + k1p] & MSK].B));
    const float *__restrict__ Sxx = p.S[0], *__restrict__ Syy = p.S[1], *__restrict__ Szz = p.S[2];
    const float *__restrict__ Sxy = p.S[3], *__restrict__ Sxz = p.S[4], *__restrict__ Syz = p.S[5];
    float X[9];
    X[0] = D4C(ci.caf, ci.cbf, Sxx[q + s1], Sxx[q], Sxx[q + 2 * s1], Sxx[q - s1]);
    X[1] = D4C(cj.cab, cj.cbb, Sxy[q], Sxy[q - j1m], Sxy[q + j1p], Sxy[q - j2m]);
    X[2] = D4C(ck.cab, ck.cbb, Sxz[q], Sxz[q - k1m], Sxz[q + k1p], Sxz[q - k2m]);
    X[3] = D4C(ci.cab, ci.cbb, Sxy[q], Sxy[q - s1], Sxy[q + s1], Sxy[q - 2 * s1]);
    X[4] = D4C(cj.caf, cj.cbf, Syy[q + j1p], Syy[q], Syy[q + j2p], Syy[q - j1m]);
    X[5] = D4C(ck.cab, ck.cbb, Syz[q], Syz[q - k1m], Syz[q + k1p], Syz[q - k2m]);
    X[6] = D4C(ci.cab, ci.cbb, Sxz[q], Sxz[q - s1], Sxz[q + s1], Sxz[q - 2 * s1]);
    X[7] = D4C(cj.cab, cj.cbb, Syz[q], Syz[q - j1m], Syz[q + j1p], Syz[q - j2m]);
    X[8] = D4C(ck.caf, ck.cbf, Szz[q + k1p], Szz[q], Szz[q + k2p], Szz[q - k1m]);
    float v[3] = { p.V[0][q], p.V[1][q], p.V[2][q] };
    if (pml) {
        const PmlCell pc = make_pml_cell(p, i, j, k);
        particle_pml<false>(p, pc, bx, by, bz, X, v);
    } else {
        v[0] += p.dt * bx * (X[0] + X[1] + X[2]);
        v[1] += p.dt * by * (X[3] + X[4] + X[5]);
        v[2] += p.dt * bz * (X[6] + X[7] + X[8]);
    }
    if (l0 & LabelTraits<LT>::REFL) { v[0] = v[1] = v[2] = 0.0f; }
    p.V[0][q] = v[0]; p.V[1][q] = v[1]; p.V[2][q] = v[2];
    if (ACC && !pml) {
        const long long qa = q - 2 * s1;
        accumulate(p, BB_MAP_VX, qa, v[0], false);
        accumulate(p, BB_MAP_VY, qa, v[1], false);
        accumulate(p, BB_MAP_VZ, qa, v[2], false);
        accumulate(p, BB_MAP_ALLV, qa, v[0] * v[0] + v[1] * v[1] + v[2] * v[2], true);
    }
}
}  // namespace direct
