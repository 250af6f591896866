// Per-cell update rules of the viscoelastic staggered-grid scheme (device functions shared by the
// direct kernels of fdtd_direct.cuh and the TMA-pipelined kernels of fdtd_tma.cuh).  The kernels
// differ only in how they fetch operands and form the nine staggered differences; everything
// after that is here.
#pragma once
#include "common.h"

#define BB_CA 1.125f
#define BB_CB (1.0f / 24.0f)
#define D4(f1, f0, f2, fm1) (BB_CA * ((f1) - (f0)) - BB_CB * ((f2) - (fm1)))
#define D4C(ca, cb, f1, f0, f2, fm1) ((ca) * ((f1) - (f0)) - (cb) * ((f2) - (fm1)))

template <typename LT> struct LabelTraits;
template <> struct LabelTraits<uint8_t> { static constexpr unsigned REFL = 0x80u, MASK = 0x7Fu; };
template <> struct LabelTraits<uint16_t> { static constexpr unsigned REFL = 0x8000u, MASK = 0x7FFFu; };

__device__ __forceinline__ MatCoef load_coef_global(const MatCoef *t, unsigned m) {
    const float4 *q = reinterpret_cast<const float4 *>(t + m);
    const float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    MatCoef r;
    r.LM = a.x; r.Mi2 = a.y; r.LMCb = a.z; r.MCb = a.w; r.a = b.x; r.cs = b.y; r.K = b.z; r.invG = b.w;
    r.tauS = c.x; r.B = c.y; r.M = c.z; r.L = c.w;
    return r;
}
__device__ __forceinline__ AxisCoef load_axis(const AxisCoef *t, int n) {
    const float4 *q = reinterpret_cast<const float4 *>(t + n);
    const float4 a = __ldg(q), b = __ldg(q + 1);
    AxisCoef r;
    r.eI = a.x; r.eH = a.y; r.pad0 = a.z; r.pad1 = a.w; r.cab = b.x; r.cbb = b.y; r.caf = b.z; r.cbf = b.w;
    return r;
}

__device__ __forceinline__ bool in_pml1(int n, int N, int P) { return n < P || n >= N - P; }
__device__ __forceinline__ bool attenuates(const MatCoef &c) { return c.LMCb != 0.0f || c.tauS != 0.0f; }
// 4/(1/g1+1/g2+1/g3+1/g4); a fluid neighbour has 1/G = +inf -> 0
__device__ __forceinline__ float rigidity4(float a, float b, float c, float d) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a + b + c + d));   // rcp(+inf) = +0; 1 ulp is far inside the 1e-4 budget
    return 4.0f * r;
}

// ------------------------------------------------------------------------------------------
// interior (non-PML) rules
// ------------------------------------------------------------------------------------------
// normal stresses + pressure accumulator; att = this material has memory variables
__device__ __forceinline__ void stress_normal_interior(const MatCoef &c, float dt, bool att, float Dxx, float Dyy, float Dzz,
                                                       float &sxx, float &syy, float &szz, float &rxx, float &ryy, float &rzz,
                                                       float &pr) {
    const float th = Dxx + Dyy + Dzz;
    pr += dt * th;
    const float o0 = Dyy + Dzz, o1 = Dxx + Dzz, o2 = Dxx + Dyy;
    const float lmth = c.LM * th;
    if (att) {
        const float rb = -c.LMCb * th;
        const float n0 = c.a * rxx + rb + c.MCb * o0;
        const float n1 = c.a * ryy + rb + c.MCb * o1;
        const float n2 = c.a * rzz + rb + c.MCb * o2;
        sxx += dt * (lmth - c.Mi2 * o0 + 0.5f * (rxx + n0));
        syy += dt * (lmth - c.Mi2 * o1 + 0.5f * (ryy + n1));
        szz += dt * (lmth - c.Mi2 * o2 + 0.5f * (rzz + n2));
        rxx = n0; ryy = n1; rzz = n2;
    } else {
        sxx += dt * (lmth - c.Mi2 * o0);
        syy += dt * (lmth - c.Mi2 * o1);
        szz += dt * (lmth - c.Mi2 * o2);
    }
}

// one shear stress on an edge with rigidity rig != 0; te = mean tau_S of the four cells
__device__ __forceinline__ void stress_shear_interior(const MatCoef &c, float dt, float rig, float te, float D, float &s, float &r) {
    if (te != 0.0f) {
        const float n = c.a * r - c.cs * (rig * te) * D;
        s += dt * (rig * (1.0f + te) * D + 0.5f * (r + n));
        r = n;
    } else {
        s += dt * (rig * D);
    }
}

// ------------------------------------------------------------------------------------------
// PML rules.  The reference formulation keeps three split parts per field and sums them.  Here a part is stored only
// where its own axis is damped; the parts of the other axes all carry the same multi-axial damping
// (mpml x the damping of the damped axes), so their sum -- the "rest" of the field, rest = f - sum(stored parts) --
// advances as one quantity:  rest' = a_r rest + b_r C sum(D of the undamped axes).  Identical to the 24-array
// formulation of the oracle up to rounding.
// ------------------------------------------------------------------------------------------
struct PmlCell {
    bool xd, jd, kd;                  // which axes are damped at this cell
    unsigned qx, qy, qz;              // index of the cell in the X / Y / Z part arrays (32 bits: checked at create)
    const AxisCoef *cI, *cJ, *cK;     // coefficient rows of the cell's i, j, k (global or shared); read only for damped axes
};

// f' = a f + b C D for half-damping e:  b = 1/(1/dt + e), a = (1/dt - e) b
__device__ __forceinline__ void pml_ab(float idt, float e, float &a, float &b) {
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(b) : "f"(idt + e));
    a = (idt - e) * b;
}

// one stored part: new value written back; old and new are accumulated for the caller.  STAGED: the old value is
// already on chip (TMA-staged box, one float per cell at `o`), otherwise it is read from the part array.
template <bool STAGED>
__device__ __forceinline__ void pml_part(const float *o, float *__restrict__ part, unsigned q, float a, float b, float CD,
                                         float &sum_old, float &sum_new) {
    const float old = STAGED ? *o : part[q];
    const float n = a * old + b * CD;
    part[q] = n;
    sum_old += old; sum_new += n;
}

__device__ __forceinline__ PmlCell make_pml_cell(const DevParams &p, int i, int j, int k) {
    PmlCell c;
    c.xd = in_pml1(i, p.n1, p.P); c.jd = in_pml1(j, p.n2, p.P); c.kd = in_pml1(k, p.n3, p.P);
    const int ipx = i < p.P ? i - p.i0 : p.nxlo + (i - p.xhi_begin);
    const int tj = j / BB_TY;
    const int jp = (tj < p.nylo ? tj : tj - p.tjhi0 + p.nylo) * BB_TY + (j - tj * BB_TY);
    const int kp = k < p.P ? k : p.zbw + (k - (p.n3 - p.P));
    c.qx = ((unsigned)ipx * p.n2 + j) * p.pitch + k;
    c.qy = ((unsigned)(i - p.i0) * p.nyrows + jp) * p.pitch + k;
    c.qz = ((unsigned)(i - p.i0) * p.n2 + j) * p.zpw + kp;
    c.cI = p.axI + i; c.cJ = p.axJ + j; c.cK = p.axK + k;
    return c;
}

// D[9] = Dxx, Dyy, Dzz, Dyx (d+_i Vy), Dxy (d+_j Vx), Dzx (d+_i Vz), Dxz (d+_k Vx), Dzy (d+_j Vz), Dyz (d+_k Vy)
// ox/oy/oz: this cell's slot in the first staged X/Y/Z part box (X and Y boxes B floats apart, Z boxes ZB floats
// apart, the two shear Z parts starting at ozs); unused when !STAGED.
template <bool STAGED>
__device__ __forceinline__ void stress_pml(const DevParams &p, const PmlCell &c, float M, float L, float rigxy, float rigxz, float rigyz,
                                           const float *D, float *s, const float *ox = nullptr, const float *oy = nullptr,
                                           const float *oz = nullptr, const float *ozs = nullptr, int B = 0, int ZB = 0) {
    const float idt = p.idt, r = p.mpml;
    float eIx = 0.f, eHx = 0.f, eIy = 0.f, eHy = 0.f, eIz = 0.f, eHz = 0.f;
    if (c.xd) { eIx = c.cI->eI; eHx = c.cI->eH; }
    if (c.jd) { eIy = c.cJ->eI; eHy = c.cJ->eH; }
    if (c.kd) { eIz = c.cK->eI; eHz = c.cK->eH; }
    const float esum = eIx + eIy + eIz;
    float so[6] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f }, sn[6] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };   // stored parts: old / new sums
    float dr[6] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };                                               // C D of the undamped axes
    float a, b;
    if (c.xd) {
        pml_ab(idt, eIx + r * (esum - eIx), a, b);
        pml_part<STAGED>(ox, p.XP[0], c.qx, a, b, M * D[0], so[0], sn[0]);
        pml_part<STAGED>(ox + B, p.XP[1], c.qx, a, b, L * D[0], so[1], sn[1]);
        pml_part<STAGED>(ox + 2 * B, p.XP[2], c.qx, a, b, L * D[0], so[2], sn[2]);
        if (rigxy != 0.0f || rigxz != 0.0f) {
            pml_ab(idt, eHx + r * (esum - eIx), a, b);
            if (rigxy != 0.0f) pml_part<STAGED>(ox + 3 * B, p.XP[3], c.qx, a, b, rigxy * D[3], so[3], sn[3]);
            if (rigxz != 0.0f) pml_part<STAGED>(ox + 4 * B, p.XP[4], c.qx, a, b, rigxz * D[5], so[4], sn[4]);
        }
    } else {
        dr[0] += M * D[0]; dr[1] += L * D[0]; dr[2] += L * D[0]; dr[3] += D[3]; dr[4] += D[5];
    }
    if (c.jd) {
        pml_ab(idt, eIy + r * (esum - eIy), a, b);
        pml_part<STAGED>(oy, p.YP[0], c.qy, a, b, L * D[1], so[0], sn[0]);
        pml_part<STAGED>(oy + B, p.YP[1], c.qy, a, b, M * D[1], so[1], sn[1]);
        pml_part<STAGED>(oy + 2 * B, p.YP[2], c.qy, a, b, L * D[1], so[2], sn[2]);
        if (rigxy != 0.0f || rigyz != 0.0f) {
            pml_ab(idt, eHy + r * (esum - eIy), a, b);
            if (rigxy != 0.0f) pml_part<STAGED>(oy + 3 * B, p.YP[3], c.qy, a, b, rigxy * D[4], so[3], sn[3]);
            if (rigyz != 0.0f) pml_part<STAGED>(oy + 4 * B, p.YP[4], c.qy, a, b, rigyz * D[7], so[5], sn[5]);
        }
    } else {
        dr[0] += L * D[1]; dr[1] += M * D[1]; dr[2] += L * D[1]; dr[3] += D[4]; dr[5] += D[7];
    }
    if (c.kd) {
        pml_ab(idt, eIz + r * (esum - eIz), a, b);
        pml_part<STAGED>(oz, p.ZP[0], c.qz, a, b, L * D[2], so[0], sn[0]);
        pml_part<STAGED>(oz + ZB, p.ZP[1], c.qz, a, b, L * D[2], so[1], sn[1]);
        pml_part<STAGED>(oz + 2 * ZB, p.ZP[2], c.qz, a, b, M * D[2], so[2], sn[2]);
        if (rigxz != 0.0f || rigyz != 0.0f) {
            pml_ab(idt, eHz + r * (esum - eIz), a, b);
            if (rigxz != 0.0f) pml_part<STAGED>(ozs, p.ZP[3], c.qz, a, b, rigxz * D[6], so[4], sn[4]);
            if (rigyz != 0.0f) pml_part<STAGED>(ozs + ZB, p.ZP[4], c.qz, a, b, rigyz * D[8], so[5], sn[5]);
        }
    } else {
        dr[0] += L * D[2]; dr[1] += L * D[2]; dr[2] += M * D[2]; dr[4] += D[6]; dr[5] += D[8];
    }
    // the parts of the undamped axes, as one quantity per field
    pml_ab(idt, r * esum, a, b);
    s[0] = sn[0] + a * (s[0] - so[0]) + b * dr[0];
    s[1] = sn[1] + a * (s[1] - so[1]) + b * dr[1];
    s[2] = sn[2] + a * (s[2] - so[2]) + b * dr[2];
    if (rigxy != 0.0f) s[3] = sn[3] + a * (s[3] - so[3]) + b * (rigxy * dr[3]);
    if (rigxz != 0.0f) s[4] = sn[4] + a * (s[4] - so[4]) + b * (rigxz * dr[4]);
    if (rigyz != 0.0f) s[5] = sn[5] + a * (s[5] - so[5]) + b * (rigyz * dr[5]);
}

// X[9] = x1 (d+_i Sxx), x2 (d-_j Sxy), x3 (d-_k Sxz), y1 (d-_i Sxy), y2 (d+_j Syy), y3 (d-_k Syz),
//        z1 (d-_i Sxz), z2 (d-_j Syz), z3 (d+_k Szz);  b = averaged 1/(rho h) of the three faces
template <bool STAGED>
__device__ __forceinline__ void particle_pml(const DevParams &p, const PmlCell &c, float bx, float by, float bz, const float *X, float *v,
                                             const float *ox = nullptr, const float *oy = nullptr, const float *oz = nullptr,
                                             int B = 0, int ZB = 0) {
    const float idt = p.idt, r = p.mpml;
    float eIx = 0.f, eHx = 0.f, eIy = 0.f, eHy = 0.f, eIz = 0.f, eHz = 0.f;
    if (c.xd) { eIx = c.cI->eI; eHx = c.cI->eH; }
    if (c.jd) { eIy = c.cJ->eI; eHy = c.cJ->eH; }
    if (c.kd) { eIz = c.cK->eI; eHz = c.cK->eH; }
    const float esum = eIx + eIy + eIz;
    float so[3] = { 0.f, 0.f, 0.f }, sn[3] = { 0.f, 0.f, 0.f }, dr[3] = { 0.f, 0.f, 0.f };
    float a, b, ah, bh;
    if (c.xd) {      // Vx sits on a half node of i, Vy and Vz on integer nodes
        pml_ab(idt, eIx + r * (esum - eIx), a, b);
        pml_ab(idt, eHx + r * (esum - eIx), ah, bh);
        pml_part<STAGED>(ox, p.XP[5], c.qx, ah, bh, bx * X[0], so[0], sn[0]);
        pml_part<STAGED>(ox + B, p.XP[6], c.qx, a, b, by * X[3], so[1], sn[1]);
        pml_part<STAGED>(ox + 2 * B, p.XP[7], c.qx, a, b, bz * X[6], so[2], sn[2]);
    } else { dr[0] += X[0]; dr[1] += X[3]; dr[2] += X[6]; }
    if (c.jd) {
        pml_ab(idt, eIy + r * (esum - eIy), a, b);
        pml_ab(idt, eHy + r * (esum - eIy), ah, bh);
        pml_part<STAGED>(oy, p.YP[5], c.qy, a, b, bx * X[1], so[0], sn[0]);
        pml_part<STAGED>(oy + B, p.YP[6], c.qy, ah, bh, by * X[4], so[1], sn[1]);
        pml_part<STAGED>(oy + 2 * B, p.YP[7], c.qy, a, b, bz * X[7], so[2], sn[2]);
    } else { dr[0] += X[1]; dr[1] += X[4]; dr[2] += X[7]; }
    if (c.kd) {
        pml_ab(idt, eIz + r * (esum - eIz), a, b);
        pml_ab(idt, eHz + r * (esum - eIz), ah, bh);
        pml_part<STAGED>(oz, p.ZP[5], c.qz, a, b, bx * X[2], so[0], sn[0]);
        pml_part<STAGED>(oz + ZB, p.ZP[6], c.qz, a, b, by * X[5], so[1], sn[1]);
        pml_part<STAGED>(oz + 2 * ZB, p.ZP[7], c.qz, ah, bh, bz * X[8], so[2], sn[2]);
    } else { dr[0] += X[2]; dr[1] += X[5]; dr[2] += X[8]; }
    pml_ab(idt, r * esum, a, b);
    v[0] = sn[0] + a * (v[0] - so[0]) + b * (bx * dr[0]);
    v[1] = sn[1] + a * (v[1] - so[1]) + b * (by * dr[1]);
    v[2] = sn[2] + a * (v[2] - so[2]) + b * (bz * dr[2]);
}

// ------------------------------------------------------------------------------------------
// RMS / peak accumulation (non-PML cells, inside the sensor window)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int acc_slot(unsigned sel, int map) { return __popc(sel & ((1u << map) - 1u)); }

__device__ __forceinline__ void accumulate(const DevParams &p, int map, unsigned qa, float v, bool squared_already) {
    if (!(p.sel_maps & (1u << map))) return;
    const long long o = (long long)acc_slot(p.sel_maps, map) * p.acc_stride + qa;
    if (p.sel_rms_peak & 1) p.acc_rms[o] += squared_already ? v : v * v;
    if (p.sel_rms_peak & 2) { if (v > p.acc_peak[o]) p.acc_peak[o] = v; }
}
