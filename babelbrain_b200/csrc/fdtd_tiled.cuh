// 2.5-D interior kernels (register queue along the slab axis i, shared-memory tile of the current
// plane for the in-plane stencil) and the box-list PML-shell kernel.
//
// Mapping: threadIdx.x -> k (contiguous axis, 128-byte aligned 32-wide tiles over the padded row),
// threadIdx.y -> j, each CTA marches `chunk` planes along i.  Lanes whose k falls in the PML
// columns (k < P or k >= n3-P) only feed the shared tile; the PML kernel owns those cells.
#pragma once
#include "fdtd_cell.cuh"

namespace tiled {
constexpr int TX = 32, TY = 8, HALO = 2;
constexpr int SW = TX + 2 * HALO;   // 36
constexpr int SH = TY + 2 * HALO;   // 12
constexpr int NHALO = SW * SH - TX * TY;  // 176 halo cells of the tile
constexpr int NT = TX * TY;

// which halo cell of the (SH x SW) tile this thread fills; soff < 0 -> none
__device__ __forceinline__ void halo_setup(int tid, int j0, int k0, const DevParams &p, int &soff, long long &goff) {
    soff = -1; goff = 0;
    if (tid >= NHALO) return;
    int r, c;
    if (tid < 4 * SW) { const int rr = tid / SW; c = tid - rr * SW; r = rr < 2 ? rr : TY + rr; }
    else { const int e = tid - 4 * SW; r = 2 + (e >> 2); const int cc = e & 3; c = cc < 2 ? cc : TX + cc; }
    const int jj = j0 - HALO + r, kk = k0 - HALO + c;
    if (jj >= 0 && jj < p.n2 && kk >= 0 && kk < p.pitch) { soff = r * SW + c; goff = (long long)jj * p.pitch + kk; }
}

#define D4(a, b, c, d) (BB_CA * ((a) - (b)) - BB_CB * ((c) - (d)))

// ------------------------------------------------------------------------------------------
// stress half-step, interior box
// ------------------------------------------------------------------------------------------
template <typename LT, bool ACC>
__global__ void __launch_bounds__(NT, 2) stress_tiled(const DevParams p, int ia, int ie, int chunk) {
    __shared__ float sV[2][3][SH * SW];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
    const int k0 = blockIdx.x * TX, j0 = p.P + blockIdx.y * TY;
    const int ic0 = ia + blockIdx.z * chunk, ic1 = min(ic0 + chunk, ie);
    const int k = k0 + tx, j = j0 + ty;
    const int jl = min(j, p.n2 - 1);  // clamp for loads of overhanging rows (results unused)
    const bool active = (k >= p.P) && (k < p.n3 - p.P) && (j < p.n2 - p.P);
    const long long col = (long long)jl * p.pitch + k;
    const long long s1 = p.plane, s2 = p.pitch;
    int hs; long long hg;
    halo_setup(tid, j0, k0, p, hs, hg);
    for (int t = tid; t < 2 * 3 * SH * SW; t += NT) (&sV[0][0][0])[t] = 0.0f;
    __syncthreads();
    const LT *lab = reinterpret_cast<const LT *>(p.lab);
    const unsigned MSK = LabelTraits<LT>::MASK;
    const float *__restrict__ Vx = p.V[0], *__restrict__ Vy = p.V[1], *__restrict__ Vz = p.V[2];
    const float dt = p.dt;
    long long q = (long long)(ic0 - p.i0 + 2) * s1 + col;  // padded index of (ic0, j, k)
    // queue prologue: Vx holds i-2..i+1, Vy/Vz hold i-1..i+2 (state before the shift of plane ic0)
    float vx_m2, vx_m1 = Vx[q - 2 * s1], vx_0 = Vx[q - s1], vx_p1 = Vx[q];
    float vy_m1, vy_0 = Vy[q - s1], vy_p1 = Vy[q], vy_p2 = Vy[q + s1];
    float vz_m1, vz_0 = Vz[q - s1], vz_p1 = Vz[q], vz_p2 = Vz[q + s1];
    const int sc = (ty + HALO) * SW + tx + HALO;
    int buf = 0;
    for (int i = ic0; i < ic1; i++, q += s1, buf ^= 1) {
        vx_m2 = vx_m1; vx_m1 = vx_0; vx_0 = vx_p1; vx_p1 = Vx[q + s1];
        vy_m1 = vy_0; vy_0 = vy_p1; vy_p1 = vy_p2; vy_p2 = Vy[q + 2 * s1];
        vz_m1 = vz_0; vz_0 = vz_p1; vz_p1 = vz_p2; vz_p2 = Vz[q + 2 * s1];
        float *bx = sV[buf][0], *by = sV[buf][1], *bz = sV[buf][2];
        bx[sc] = vx_0; by[sc] = vy_0; bz[sc] = vz_0;
        if (hs >= 0) {
            const long long g = (long long)(i - p.i0 + 2) * s1 + hg;
            bx[hs] = Vx[g]; by[hs] = Vy[g]; bz[hs] = Vz[g];
        }
        __syncthreads();
        if (!active) continue;
        const unsigned l0 = lab[q];
        const unsigned m = l0 & MSK;
        const bool refl = (l0 & LabelTraits<LT>::REFL) != 0;
        const MatRow r = load_mat(p.mat, m);
        const float Dxx = D4(vx_0, vx_m1, vx_p1, vx_m2);
        const float Dyy = D4(by[sc], by[sc - SW], by[sc + SW], by[sc - 2 * SW]);
        const float Dzz = D4(bz[sc], bz[sc - 1], bz[sc + 1], bz[sc - 2]);
        const float th = Dxx + Dyy + Dzz;
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
        sv[3] = sv[4] = sv[5] = 0.0f;
        if (r.G != 0.0f) {  // fluid cells have zero edge rigidity on all three edges
            const unsigned mi = lab[q + s1] & MSK, mj = lab[q + s2] & MSK, mk = lab[q + 1] & MSK;
            const unsigned mij = lab[q + s1 + s2] & MSK, mik = lab[q + s1 + 1] & MSK, mjk = lab[q + s2 + 1] & MSK;
            const float gi = mat_G(p.mat, mi), gj = mat_G(p.mat, mj), gk = mat_G(p.mat, mk);
            const float rig[3] = { harm4(r.G, gi, gj, mat_G(p.mat, mij)), harm4(r.G, gi, gk, mat_G(p.mat, mik)),
                                   harm4(r.G, gj, gk, mat_G(p.mat, mjk)) };
#pragma unroll
            for (int c = 0; c < 3; c++) {
                if (rig[c] == 0.0f) { if (ACC) sv[3 + c] = p.S[3 + c][q]; continue; }
                float D, tsum;
                if (c == 0) {
                    D = D4(vy_p1, vy_0, vy_p2, vy_m1) + D4(bx[sc + SW], bx[sc], bx[sc + 2 * SW], bx[sc - SW]);
                    tsum = r.tauS + mat_tauS(p.mat, mi) + mat_tauS(p.mat, mj) + mat_tauS(p.mat, mij);
                } else if (c == 1) {
                    D = D4(vz_p1, vz_0, vz_p2, vz_m1) + D4(bx[sc + 1], bx[sc], bx[sc + 2], bx[sc - 1]);
                    tsum = r.tauS + mat_tauS(p.mat, mi) + mat_tauS(p.mat, mk) + mat_tauS(p.mat, mik);
                } else {
                    D = D4(bz[sc + SW], bz[sc], bz[sc + 2 * SW], bz[sc - SW]) + D4(by[sc + 1], by[sc], by[sc + 2], by[sc - 1]);
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
        } else if (ACC && (p.sel_maps & 0x380u)) {
            sv[3] = p.S[3][q]; sv[4] = p.S[4][q]; sv[5] = p.S[5][q];
        }
        if (ACC) {
            const long long qa = q - 2 * s1;
#pragma unroll
            for (int c = 0; c < 6; c++) accumulate(p, BB_MAP_SXX + c, qa, sv[c], false);
            accumulate(p, BB_MAP_PRESSURE, qa, -r.K * pr, false);
        }
    }
}

// ------------------------------------------------------------------------------------------
// particle half-step, interior box
// ------------------------------------------------------------------------------------------
template <typename LT, bool ACC>
__global__ void __launch_bounds__(NT, 2) particle_tiled(const DevParams p, int ia, int ie, int chunk) {
    // in-plane operands of the current plane: Sxy, Sxz, Syy, Syz, Szz
    __shared__ float sS[2][5][SH * SW];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
    const int k0 = blockIdx.x * TX, j0 = p.P + blockIdx.y * TY;
    const int ic0 = ia + blockIdx.z * chunk, ic1 = min(ic0 + chunk, ie);
    const int k = k0 + tx, j = j0 + ty;
    const int jl = min(j, p.n2 - 1);
    const bool active = (k >= p.P) && (k < p.n3 - p.P) && (j < p.n2 - p.P);
    const long long col = (long long)jl * p.pitch + k;
    const long long s1 = p.plane, s2 = p.pitch;
    int hs; long long hg;
    halo_setup(tid, j0, k0, p, hs, hg);
    for (int t = tid; t < 2 * 5 * SH * SW; t += NT) (&sS[0][0][0])[t] = 0.0f;
    __syncthreads();
    const LT *lab = reinterpret_cast<const LT *>(p.lab);
    const unsigned MSK = LabelTraits<LT>::MASK;
    const float *__restrict__ Sxx = p.S[0], *__restrict__ Syy = p.S[1], *__restrict__ Szz = p.S[2];
    const float *__restrict__ Sxy = p.S[3], *__restrict__ Sxz = p.S[4], *__restrict__ Syz = p.S[5];
    const float dt = p.dt;
    long long q = (long long)(ic0 - p.i0 + 2) * s1 + col;
    // Sxx holds i-1..i+2 ; Sxy, Sxz hold i-2..i+1
    float xx_m1, xx_0 = Sxx[q - s1], xx_p1 = Sxx[q], xx_p2 = Sxx[q + s1];
    float xy_m2, xy_m1 = Sxy[q - 2 * s1], xy_0 = Sxy[q - s1], xy_p1 = Sxy[q];
    float xz_m2, xz_m1 = Sxz[q - 2 * s1], xz_0 = Sxz[q - s1], xz_p1 = Sxz[q];
    const int sc = (ty + HALO) * SW + tx + HALO;
    int buf = 0;
    for (int i = ic0; i < ic1; i++, q += s1, buf ^= 1) {
        xx_m1 = xx_0; xx_0 = xx_p1; xx_p1 = xx_p2; xx_p2 = Sxx[q + 2 * s1];
        xy_m2 = xy_m1; xy_m1 = xy_0; xy_0 = xy_p1; xy_p1 = Sxy[q + s1];
        xz_m2 = xz_m1; xz_m1 = xz_0; xz_0 = xz_p1; xz_p1 = Sxz[q + s1];
        float *bxy = sS[buf][0], *bxz = sS[buf][1], *byy = sS[buf][2], *byz = sS[buf][3], *bzz = sS[buf][4];
        bxy[sc] = xy_0; bxz[sc] = xz_0; byy[sc] = Syy[q]; byz[sc] = Syz[q]; bzz[sc] = Szz[q];
        if (hs >= 0) {
            const long long g = (long long)(i - p.i0 + 2) * s1 + hg;
            bxy[hs] = Sxy[g]; bxz[hs] = Sxz[g]; byy[hs] = Syy[g]; byz[hs] = Syz[g]; bzz[hs] = Szz[g];
        }
        __syncthreads();
        if (!active) continue;
        const unsigned l0 = lab[q];
        const float b0 = mat_B(p.mat, l0 & MSK);
        const float bx = 0.5f * (b0 + mat_B(p.mat, lab[q + s1] & MSK));
        const float by = 0.5f * (b0 + mat_B(p.mat, lab[q + s2] & MSK));
        const float bz = 0.5f * (b0 + mat_B(p.mat, lab[q + 1] & MSK));
        const float x1 = D4(xx_p1, xx_0, xx_p2, xx_m1);
        const float x2 = D4(bxy[sc], bxy[sc - SW], bxy[sc + SW], bxy[sc - 2 * SW]);
        const float x3 = D4(bxz[sc], bxz[sc - 1], bxz[sc + 1], bxz[sc - 2]);
        const float y1 = D4(xy_0, xy_m1, xy_p1, xy_m2);
        const float y2 = D4(byy[sc + SW], byy[sc], byy[sc + 2 * SW], byy[sc - SW]);
        const float y3 = D4(byz[sc], byz[sc - 1], byz[sc + 1], byz[sc - 2]);
        const float z1 = D4(xz_0, xz_m1, xz_p1, xz_m2);
        const float z2 = D4(byz[sc], byz[sc - SW], byz[sc + SW], byz[sc - 2 * SW]);
        const float z3 = D4(bzz[sc + 1], bzz[sc], bzz[sc + 2], bzz[sc - 1]);
        float vx = p.V[0][q] + dt * bx * (x1 + x2 + x3);
        float vy = p.V[1][q] + dt * by * (y1 + y2 + y3);
        float vz = p.V[2][q] + dt * bz * (z1 + z2 + z3);
        if (l0 & LabelTraits<LT>::REFL) { vx = vy = vz = 0.0f; }
        p.V[0][q] = vx; p.V[1][q] = vy; p.V[2][q] = vz;
        if (ACC) {
            const long long qa = q - 2 * s1;
            accumulate(p, BB_MAP_VX, qa, vx, false);
            accumulate(p, BB_MAP_VY, qa, vy, false);
            accumulate(p, BB_MAP_VZ, qa, vz, false);
            accumulate(p, BB_MAP_ALLV, qa, vx * vx + vy * vy + vz * vz, true);
        }
    }
}

// ------------------------------------------------------------------------------------------
// PML shell: up to 6 boxes in one launch, one thread per cell
// ------------------------------------------------------------------------------------------
struct PmlBox { int i0, i1, j0, j1, k0, k1; int nbj, nbk; int bx_shift; int first_block; };
struct PmlBoxes { PmlBox b[6]; int n; int total_blocks; };

template <typename LT, bool STRESS>
__global__ void __launch_bounds__(256) pml_boxes_kernel(const DevParams p, const PmlBoxes boxes) {
    int bi = 0;
#pragma unroll
    for (int n = 1; n < 6; n++) if (n < boxes.n && (int)blockIdx.x >= boxes.b[n].first_block) bi = n;
    const PmlBox &b = boxes.b[bi];
    int r = blockIdx.x - b.first_block;
    const int per_plane = b.nbj * b.nbk;
    const int ip = r / per_plane; r -= ip * per_plane;
    const int bj = r / b.nbk, bk = r - bj * b.nbk;
    const int bxs = b.bx_shift;                    // 5 -> 32x8 threads, 4 -> 16x16
    const int tx = threadIdx.x & ((1 << bxs) - 1), ty = threadIdx.x >> bxs;
    const int k = b.k0 + (bk << bxs) + tx, j = b.j0 + bj * (256 >> bxs) + ty, i = b.i0 + ip;
    if (k >= b.k1 || j >= b.j1) return;
    const long long q = ((long long)(i - p.i0 + 2) * p.n2 + j) * p.pitch + k;
    if (STRESS) stress_cell_pml<LT>(p, i, j, k, q);
    else particle_cell_pml<LT>(p, i, j, k, q);
}

static inline void add_box(PmlBoxes &B, int i0, int i1, int j0, int j1, int k0, int k1) {
    if (i1 <= i0 || j1 <= j0 || k1 <= k0) return;
    PmlBox &b = B.b[B.n++];
    b.i0 = i0; b.i1 = i1; b.j0 = j0; b.j1 = j1; b.k0 = k0; b.k1 = k1;
    b.bx_shift = (k1 - k0) <= 16 ? 4 : 5;
    const int bx = 1 << b.bx_shift, by = 256 >> b.bx_shift;
    b.nbk = (k1 - k0 + bx - 1) / bx;
    b.nbj = (j1 - j0 + by - 1) / by;
    b.first_block = B.total_blocks;
    B.total_blocks += (i1 - i0) * b.nbj * b.nbk;
}

static inline PmlBoxes make_pml_boxes(const DevParams &p, int ib, int ie) {
    PmlBoxes B;
    B.n = 0; B.total_blocks = 0;
    const int P = p.P;
    add_box(B, ib, std::min(ie, P), 0, p.n2, 0, p.n3);
    add_box(B, std::max(ib, p.n1 - P), ie, 0, p.n2, 0, p.n3);
    const int im0 = std::max(ib, P), im1 = std::min(ie, p.n1 - P);
    add_box(B, im0, im1, 0, P, 0, p.n3);
    add_box(B, im0, im1, p.n2 - P, p.n2, 0, p.n3);
    add_box(B, im0, im1, P, p.n2 - P, 0, P);
    add_box(B, im0, im1, P, p.n2 - P, p.n3 - P, p.n3);
    return B;
}
}  // namespace tiled

// launch one half-step over the owned planes [ib, ie): interior box (tiled) + PML shell (boxes)
template <typename LT, typename FB, typename FE>
static int launch_tiled(const DevParams &p, bool stress, bool acc, int ib, int ie, cudaStream_t st, FB tbegin, FE tend) {
    using namespace tiled;
    const int ia = std::max(ib, p.P), iz = std::min(ie, p.n1 - p.P);
    if (iz > ia) {
        const int nplanes = iz - ia;
        int chunk = nplanes >= 64 ? 32 : (nplanes >= 16 ? 16 : nplanes);
        const dim3 blk(TX, TY, 1), grid(p.pitch / TX, (p.n2 - 2 * p.P + TY - 1) / TY, (nplanes + chunk - 1) / chunk);
        tbegin(0);
        if (stress) {
            if (acc) stress_tiled<LT, true><<<grid, blk, 0, st>>>(p, ia, iz, chunk);
            else stress_tiled<LT, false><<<grid, blk, 0, st>>>(p, ia, iz, chunk);
        } else {
            if (acc) particle_tiled<LT, true><<<grid, blk, 0, st>>>(p, ia, iz, chunk);
            else particle_tiled<LT, false><<<grid, blk, 0, st>>>(p, ia, iz, chunk);
        }
        tend();
        BB_CUDA(cudaGetLastError());
    }
    const PmlBoxes B = make_pml_boxes(p, ib, ie);
    if (B.total_blocks > 0) {
        tbegin(1);
        if (stress) pml_boxes_kernel<LT, true><<<B.total_blocks, 256, 0, st>>>(p, B);
        else pml_boxes_kernel<LT, false><<<B.total_blocks, 256, 0, st>>>(p, B);
        tend();
        BB_CUDA(cudaGetLastError());
    }
    return BB_OK;
}
