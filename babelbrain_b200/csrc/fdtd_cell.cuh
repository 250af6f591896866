// Per-cell update rules of the viscoelastic staggered-grid scheme (device functions).
// These are the reference formulation every tiled kernel in fdtd_kernels.cuh must reproduce:
// the PML-shell kernels call them directly, the interior kernels inline the same arithmetic on
// register/shared-memory operands.
#pragma once
#include "common.h"

#define BB_CA 1.125f
#define BB_CB (1.0f / 24.0f)

template <typename LT> struct LabelTraits;
template <> struct LabelTraits<uint8_t> { static constexpr unsigned REFL = 0x80u, MASK = 0x7Fu; };
template <> struct LabelTraits<uint16_t> { static constexpr unsigned REFL = 0x8000u, MASK = 0x7FFFu; };

__device__ __forceinline__ MatRow load_mat(const MatRow *t, unsigned m) {
    const float4 *q = reinterpret_cast<const float4 *>(t + m);
    float4 a = __ldg(q), b = __ldg(q + 1);
    MatRow r;
    r.M = a.x; r.G = a.y; r.L = a.z; r.B = a.w; r.tauL = b.x; r.tauS = b.y; r.ots = b.z; r.K = b.w;
    return r;
}
__device__ __forceinline__ float mat_G(const MatRow *t, unsigned m) { return __ldg(&t[m].G); }
__device__ __forceinline__ float mat_tauS(const MatRow *t, unsigned m) { return __ldg(&t[m].tauS); }
__device__ __forceinline__ float mat_B(const MatRow *t, unsigned m) { return __ldg(&t[m].B); }

__device__ __forceinline__ float harm4(float g1, float g2, float g3, float g4) {
    if (g1 * g2 * g3 * g4 == 0.0f) return 0.0f;
    return 4.0f / (1.0f / g1 + 1.0f / g2 + 1.0f / g3 + 1.0f / g4);
}

__device__ __forceinline__ bool in_pml1(int n, int N, int P) { return n < P || n >= N - P; }

// PML coefficients at integer and half nodes of one axis (tables of P+1 entries each)
__device__ __forceinline__ void coef_int(const DevParams &p, int n, int N, float &a, float &b) {
    int d = 0;
    if (n < p.P) d = p.P - n; else if (n >= N - p.P) d = n - (N - p.P - 1);
    a = __ldg(p.pml + d); b = __ldg(p.pml + (p.P + 1) + d);
}
__device__ __forceinline__ void coef_half(const DevParams &p, int n, int N, float &a, float &b) {
    const int P1 = p.P + 1;
    if (n < p.P) { int d = p.P - 1 - n; a = __ldg(p.pml + 2 * P1 + d); b = __ldg(p.pml + 3 * P1 + d); }
    else if (n >= N - p.P) { int d = n - (N - p.P - 1); a = __ldg(p.pml + 2 * P1 + d); b = __ldg(p.pml + 3 * P1 + d); }
    else { a = __ldg(p.pml); b = __ldg(p.pml + P1); }
}

// staggered differences with the domain-edge rules (n = global index along the axis)
__device__ __forceinline__ float dbwd_e(const float *f, long long q, long long st, int n, int N) {
    if (n > 1 && n < N - 1) return BB_CA * (f[q] - f[q - st]) - BB_CB * (f[q + st] - f[q - 2 * st]);
    if (n > 0) return f[q] - f[q - st];
    return 0.0f;
}
__device__ __forceinline__ float dfwd_e(const float *f, long long q, long long st, int n, int N) {
    if (n > 0 && n < N - 2) return BB_CA * (f[q + st] - f[q]) - BB_CB * (f[q + 2 * st] - f[q - st]);
    if (n < N - 1) return f[q + st] - f[q];
    return 0.0f;
}
// interior: always 4th order
__device__ __forceinline__ float dbwd4(const float *f, long long q, long long st) {
    return BB_CA * (f[q] - f[q - st]) - BB_CB * (f[q + st] - f[q - 2 * st]);
}
__device__ __forceinline__ float dfwd4(const float *f, long long q, long long st) {
    return BB_CA * (f[q + st] - f[q]) - BB_CB * (f[q + 2 * st] - f[q - st]);
}

// compact index of a PML-shell cell of this slab (i global)
__device__ __forceinline__ long long pml_index(const DevParams &p, int i, int j, int k) {
    if (i < p.ilo_end) return ((long long)(i - p.i0) * p.n2 + j) * p.n3 + k;
    if (i >= p.ihi_begin) return p.off[1] + ((long long)(i - p.ihi_begin) * p.n2 + j) * p.n3 + k;
    const int im = i - p.ilo_end;
    if (j < p.P) return p.off[2] + ((long long)im * p.P + j) * p.n3 + k;
    if (j >= p.n2 - p.P) return p.off[3] + ((long long)im * p.P + (j - (p.n2 - p.P))) * p.n3 + k;
    const int n2m = p.n2 - 2 * p.P;
    if (k < p.P) return p.off[4] + ((long long)im * n2m + (j - p.P)) * p.P + k;
    return p.off[5] + ((long long)im * n2m + (j - p.P)) * p.P + (k - (p.n3 - p.P));
}

__device__ __forceinline__ float split_upd(float *arr, long long q, float a, float b, float C, float D) {
    const float v = a * (arr[q] * b + C * D);
    arr[q] = v;
    return v;
}

__device__ __forceinline__ int acc_slot(unsigned sel, int map) { return __popc(sel & ((1u << map) - 1u)); }

__device__ __forceinline__ void accumulate(const DevParams &p, int map, long long qa, float v, bool squared_already) {
    if (!(p.sel_maps & (1u << map))) return;
    const long long o = (long long)acc_slot(p.sel_maps, map) * p.acc_stride + qa;
    if (p.sel_rms_peak & 1) p.acc_rms[o] += squared_already ? v : v * v;
    if (p.sel_rms_peak & 2) { if (v > p.acc_peak[o]) p.acc_peak[o] = v; }
}

// ------------------------------------------------------------------------------------------
// stress half-step, PML-shell cell (split field); i global, q = padded linear index
// ------------------------------------------------------------------------------------------
template <typename LT>
__device__ __forceinline__ void stress_cell_pml(const DevParams &p, int i, int j, int k, long long q) {
    if (!(i < p.n1 - 1 && j < p.n2 - 1 && k < p.n3 - 1)) return;
    const LT *lab = reinterpret_cast<const LT *>(p.lab);
    const unsigned MSK = LabelTraits<LT>::MASK;
    const long long s1 = p.plane, s2 = p.pitch;
    const unsigned m = lab[q] & MSK, mi = lab[q + s1] & MSK, mj = lab[q + s2] & MSK, mk = lab[q + 1] & MSK;
    const unsigned mij = lab[q + s1 + s2] & MSK, mik = lab[q + s1 + 1] & MSK, mjk = lab[q + s2 + 1] & MSK;
    const float g0 = mat_G(p.mat, m), gi = mat_G(p.mat, mi), gj = mat_G(p.mat, mj), gk = mat_G(p.mat, mk);
    const float rigxy = harm4(g0, gi, gj, mat_G(p.mat, mij));
    const float rigxz = harm4(g0, gi, gk, mat_G(p.mat, mik));
    const float rigyz = harm4(g0, gj, gk, mat_G(p.mat, mjk));
    const float *Vx = p.V[0], *Vy = p.V[1], *Vz = p.V[2];
    const float Dxx = dbwd_e(Vx, q, s1, i, p.n1), Dyy = dbwd_e(Vy, q, s2, j, p.n2), Dzz = dbwd_e(Vz, q, 1, k, p.n3);
    const long long c = pml_index(p, i, j, k);
    const float M = __ldg(&p.mat[m].M), L = __ldg(&p.mat[m].L);
    float a, b;
    coef_int(p, i, p.n1, a, b);
    float xx = split_upd(p.sp[SP_SXX_X], c, a, b, M, Dxx);
    float yy = split_upd(p.sp[SP_SYY_X], c, a, b, L, Dxx);
    float zz = split_upd(p.sp[SP_SZZ_X], c, a, b, L, Dxx);
    coef_int(p, j, p.n2, a, b);
    xx += split_upd(p.sp[SP_SXX_Y], c, a, b, L, Dyy);
    yy += split_upd(p.sp[SP_SYY_Y], c, a, b, M, Dyy);
    zz += split_upd(p.sp[SP_SZZ_Y], c, a, b, L, Dyy);
    coef_int(p, k, p.n3, a, b);
    xx += split_upd(p.sp[SP_SXX_Z], c, a, b, L, Dzz);
    yy += split_upd(p.sp[SP_SYY_Z], c, a, b, L, Dzz);
    zz += split_upd(p.sp[SP_SZZ_Z], c, a, b, M, Dzz);
    float ai, bi, aj, bj, ak, bk;
    coef_half(p, i, p.n1, ai, bi); coef_half(p, j, p.n2, aj, bj); coef_half(p, k, p.n3, ak, bk);
    float xy = split_upd(p.sp[SP_SXY_X], c, ai, bi, rigxy, dfwd_e(Vy, q, s1, i, p.n1))
             + split_upd(p.sp[SP_SXY_Y], c, aj, bj, rigxy, dfwd_e(Vx, q, s2, j, p.n2));
    float xz = split_upd(p.sp[SP_SXZ_X], c, ai, bi, rigxz, dfwd_e(Vz, q, s1, i, p.n1))
             + split_upd(p.sp[SP_SXZ_Z], c, ak, bk, rigxz, dfwd_e(Vx, q, 1, k, p.n3));
    float yz = split_upd(p.sp[SP_SYZ_Y], c, aj, bj, rigyz, dfwd_e(Vz, q, s2, j, p.n2))
             + split_upd(p.sp[SP_SYZ_Z], c, ak, bk, rigyz, dfwd_e(Vy, q, 1, k, p.n3));
    if (lab[q] & LabelTraits<LT>::REFL) { xx = yy = zz = xy = xz = yz = 0.0f; }
    p.S[0][q] = xx; p.S[1][q] = yy; p.S[2][q] = zz; p.S[3][q] = xy; p.S[4][q] = xz; p.S[5][q] = yz;
}

// ------------------------------------------------------------------------------------------
// stress half-step, interior cell (viscoelastic, 4th order), straight from global memory
// ------------------------------------------------------------------------------------------
template <typename LT, bool ACC>
__device__ __forceinline__ void stress_cell_interior(const DevParams &p, int i, int j, int k, long long q) {
    const LT *lab = reinterpret_cast<const LT *>(p.lab);
    const unsigned MSK = LabelTraits<LT>::MASK;
    const long long s1 = p.plane, s2 = p.pitch;
    const unsigned l0 = lab[q];
    const unsigned m = l0 & MSK, mi = lab[q + s1] & MSK, mj = lab[q + s2] & MSK, mk = lab[q + 1] & MSK;
    const unsigned mij = lab[q + s1 + s2] & MSK, mik = lab[q + s1 + 1] & MSK, mjk = lab[q + s2 + 1] & MSK;
    const MatRow r = load_mat(p.mat, m);
    const float *Vx = p.V[0], *Vy = p.V[1], *Vz = p.V[2];
    const float dt = p.dt;
    const float Dxx = dbwd4(Vx, q, s1), Dyy = dbwd4(Vy, q, s2), Dzz = dbwd4(Vz, q, 1);
    const float th = Dxx + Dyy + Dzz;
    const bool refl = (l0 & LabelTraits<LT>::REFL) != 0;
    float pr = p.Pr[q] + dt * th;
    if (refl) pr = 0.0f;
    p.Pr[q] = pr;
    const float LM = r.M * (1.0f + r.tauL), Mi2 = 2.0f * r.G * (1.0f + r.tauS);
    const bool att = (r.tauL != 0.0f) || (r.tauS != 0.0f);
    const float LMC = dt * r.M * (r.tauL * r.ots), MC = dt * 2.0f * r.G * (r.tauS * r.ots);
    const float den = 1.0f + dt * 0.5f * r.ots, num = 1.0f - dt * 0.5f * r.ots;
    const float oth[3] = { Dyy + Dzz, Dxx + Dzz, Dxx + Dyy };
    float sv[6];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float s = p.S[c][q];
        if (att) {
            const float R = p.R[c][q];
            const float NextR = (num * R - LMC * th + MC * oth[c]) / den;
            s += dt * (LM * th - Mi2 * oth[c] + 0.5f * (R + NextR));
            p.R[c][q] = NextR;
        } else {
            s += dt * (LM * th - Mi2 * oth[c]);
        }
        if (refl) s = 0.0f;
        p.S[c][q] = s;
        sv[c] = s;
    }
    const float g0 = r.G, gi = mat_G(p.mat, mi), gj = mat_G(p.mat, mj), gk = mat_G(p.mat, mk);
    const float rig[3] = { harm4(g0, gi, gj, mat_G(p.mat, mij)), harm4(g0, gi, gk, mat_G(p.mat, mik)),
                           harm4(g0, gj, gk, mat_G(p.mat, mjk)) };
#pragma unroll
    for (int c = 0; c < 3; c++) {
        sv[3 + c] = 0.0f;
        if (rig[c] == 0.0f) { if (ACC) sv[3 + c] = p.S[3 + c][q]; continue; }
        float D, tsum;
        if (c == 0) {
            D = dfwd4(Vy, q, s1) + dfwd4(Vx, q, s2);
            tsum = r.tauS + mat_tauS(p.mat, mi) + mat_tauS(p.mat, mj) + mat_tauS(p.mat, mij);
        } else if (c == 1) {
            D = dfwd4(Vz, q, s1) + dfwd4(Vx, q, 1);
            tsum = r.tauS + mat_tauS(p.mat, mi) + mat_tauS(p.mat, mk) + mat_tauS(p.mat, mik);
        } else {
            D = dfwd4(Vz, q, s2) + dfwd4(Vy, q, 1);
            tsum = r.tauS + mat_tauS(p.mat, mj) + mat_tauS(p.mat, mk) + mat_tauS(p.mat, mjk);
        }
        const float te = 0.25f * tsum;
        float s = p.S[3 + c][q];
        if (te != 0.0f) {
            const float R = p.R[3 + c][q];
            const float NextR = (num * R - dt * (rig[c] * (te * r.ots)) * D) / den;
            s += dt * (rig[c] * (1.0f + te) * D + 0.5f * (R + NextR));
            p.R[3 + c][q] = NextR;
        } else {
            s += dt * (rig[c] * (1.0f + te) * D);
        }
        if (refl) s = 0.0f;
        p.S[3 + c][q] = s;
        sv[3 + c] = s;
    }
    if (ACC) {
        const long long qa = q - 2 * p.plane;  // owned planes start at local plane 2
#pragma unroll
        for (int c = 0; c < 6; c++) accumulate(p, BB_MAP_SXX + c, qa, sv[c], false);
        accumulate(p, BB_MAP_PRESSURE, qa, -r.K * pr, false);
    }
}

// ------------------------------------------------------------------------------------------
// particle half-step
// ------------------------------------------------------------------------------------------
template <typename LT>
__device__ __forceinline__ void particle_cell_pml(const DevParams &p, int i, int j, int k, long long q) {
    if (!(i < p.n1 - 1 && j < p.n2 - 1 && k < p.n3 - 1)) return;
    const LT *lab = reinterpret_cast<const LT *>(p.lab);
    const unsigned MSK = LabelTraits<LT>::MASK;
    const long long s1 = p.plane, s2 = p.pitch;
    const float b0 = mat_B(p.mat, lab[q] & MSK);
    const float bx = 0.5f * (b0 + mat_B(p.mat, lab[q + s1] & MSK));
    const float by = 0.5f * (b0 + mat_B(p.mat, lab[q + s2] & MSK));
    const float bz = 0.5f * (b0 + mat_B(p.mat, lab[q + 1] & MSK));
    const float *Sxx = p.S[0], *Syy = p.S[1], *Szz = p.S[2], *Sxy = p.S[3], *Sxz = p.S[4], *Syz = p.S[5];
    const float x1 = dfwd_e(Sxx, q, s1, i, p.n1), x2 = dbwd_e(Sxy, q, s2, j, p.n2), x3 = dbwd_e(Sxz, q, 1, k, p.n3);
    const float y1 = dbwd_e(Sxy, q, s1, i, p.n1), y2 = dfwd_e(Syy, q, s2, j, p.n2), y3 = dbwd_e(Syz, q, 1, k, p.n3);
    const float z1 = dbwd_e(Sxz, q, s1, i, p.n1), z2 = dbwd_e(Syz, q, s2, j, p.n2), z3 = dfwd_e(Szz, q, 1, k, p.n3);
    float ai, bi, aj, bj, ak, bk, hi, gi, hj, gj, hk, gk;
    coef_int(p, i, p.n1, ai, bi); coef_int(p, j, p.n2, aj, bj); coef_int(p, k, p.n3, ak, bk);
    coef_half(p, i, p.n1, hi, gi); coef_half(p, j, p.n2, hj, gj); coef_half(p, k, p.n3, hk, gk);
    const long long c = pml_index(p, i, j, k);
    float vx = split_upd(p.sp[SP_VX_X], c, hi, gi, bx, x1) + split_upd(p.sp[SP_VX_Y], c, aj, bj, bx, x2)
             + split_upd(p.sp[SP_VX_Z], c, ak, bk, bx, x3);
    float vy = split_upd(p.sp[SP_VY_X], c, ai, bi, by, y1) + split_upd(p.sp[SP_VY_Y], c, hj, gj, by, y2)
             + split_upd(p.sp[SP_VY_Z], c, ak, bk, by, y3);
    float vz = split_upd(p.sp[SP_VZ_X], c, ai, bi, bz, z1) + split_upd(p.sp[SP_VZ_Y], c, aj, bj, bz, z2)
             + split_upd(p.sp[SP_VZ_Z], c, hk, gk, bz, z3);
    if (lab[q] & LabelTraits<LT>::REFL) { vx = vy = vz = 0.0f; }
    p.V[0][q] = vx; p.V[1][q] = vy; p.V[2][q] = vz;
}

template <typename LT, bool ACC>
__device__ __forceinline__ void particle_cell_interior(const DevParams &p, int i, int j, int k, long long q) {
    const LT *lab = reinterpret_cast<const LT *>(p.lab);
    const unsigned MSK = LabelTraits<LT>::MASK;
    const long long s1 = p.plane, s2 = p.pitch;
    const unsigned l0 = lab[q];
    const float b0 = mat_B(p.mat, l0 & MSK);
    const float bx = 0.5f * (b0 + mat_B(p.mat, lab[q + s1] & MSK));
    const float by = 0.5f * (b0 + mat_B(p.mat, lab[q + s2] & MSK));
    const float bz = 0.5f * (b0 + mat_B(p.mat, lab[q + 1] & MSK));
    const float *Sxx = p.S[0], *Syy = p.S[1], *Szz = p.S[2], *Sxy = p.S[3], *Sxz = p.S[4], *Syz = p.S[5];
    const float dt = p.dt;
    float vx = p.V[0][q] + dt * bx * (dfwd4(Sxx, q, s1) + dbwd4(Sxy, q, s2) + dbwd4(Sxz, q, 1));
    float vy = p.V[1][q] + dt * by * (dbwd4(Sxy, q, s1) + dfwd4(Syy, q, s2) + dbwd4(Syz, q, 1));
    float vz = p.V[2][q] + dt * bz * (dbwd4(Sxz, q, s1) + dbwd4(Syz, q, s2) + dfwd4(Szz, q, 1));
    if (l0 & LabelTraits<LT>::REFL) { vx = vy = vz = 0.0f; }
    p.V[0][q] = vx; p.V[1][q] = vy; p.V[2][q] = vz;
    if (ACC) {
        const long long qa = q - 2 * p.plane;
        accumulate(p, BB_MAP_VX, qa, vx, false);
        accumulate(p, BB_MAP_VY, qa, vy, false);
        accumulate(p, BB_MAP_VZ, qa, vz, false);
        accumulate(p, BB_MAP_ALLV, qa, vx * vx + vy * vy + vz * vz, true);
    }
}
