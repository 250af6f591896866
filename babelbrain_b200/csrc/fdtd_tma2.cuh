// kernel_variant = 2: the half-step kernels with TWO cells per thread.
//
// Same tiles, same TMA descriptors, same two mbarrier rings and the same two producer warps as fdtd_tma.cuh
// (kernel_variant 0); what changes is the consumer side.  A CTA has 8 consumer warps instead of 16, and a thread owns the
// two k-adjacent cells (ty, 2 lane) and (ty, 2 lane + 1) of the 8 x 64 tile -- on plain tiles a warp is one whole
// 256-byte tile row.  ncu on variant 0 (profiles/r1_ncu_stress_particle_summary.txt) shows kernels bound by instruction
// issue and latency, not by DRAM: ~280 warp instructions per cell and plane, of which ~100 (two mbarrier waits, ring
// bookkeeping, flag tests, register-queue moves, address forming) do not depend on the number of cells a thread owns,
// and the rest is dominated by 32-bit shared loads and 32-bit global stores.  With a cell pair per thread
//   * the per-thread overhead is paid once per two cells,
//   * every operand of the pair is one 64-bit access: LDS.64 from the halo / point boxes, STG.64 to the field arrays
//     (a warp stores 256 contiguous bytes per field and plane),
//   * the k-stencils of the two cells share their taps (five values instead of eight),
//   * each thread carries two independent dependency chains, and the 320-thread CTA may use up to 204 registers,
//     so ptxas hoists the loads of a plane ahead of its arithmetic.
// Fields a cell does not change are written back with the value staged by TMA (each cell is owned by exactly one
// CTA of a launch, so that value is the current one); the cell arithmetic itself is fdtd_cell.cuh, unchanged.
#pragma once
#include "fdtd_tma.cuh"

namespace tma {
constexpr int NCW2 = TY;                 // consumer warps: one per tile row on plain tiles
constexpr int NT2 = NCW2 * 32;           // 256 consumer threads, two cells each
constexpr int NTB2 = NT2 + 64;           // + two producer warps (halo ring, point ring)
constexpr int BOXF = TX * TY;            // floats per point box
static_assert(TX == 64 && (SW % 2) == 0 && (HK % 2) == 0, "cell pairs must be 8-byte aligned in the staged boxes");

__device__ __forceinline__ void consumer_bar2() { asm volatile("bar.sync 1, %0;" ::"n"(NT2) : "memory"); }
__device__ __forceinline__ float2 ld2(const float *p) { return *reinterpret_cast<const float2 *>(p); }
__device__ __forceinline__ void st2(float *p, float a, float b) { *reinterpret_cast<float2 *>(p) = make_float2(a, b); }

// thread -> cell pair (ty, tx), tx even.  Plain tiles: warp = row.  Tiles holding k-PML columns: the pairs that hold PML
// columns of all rows are enumerated first, then the interior pairs, so that all but one warp run a single path.
template <int ROWS = TY>
__device__ __forceinline__ bool map_pair(const DevParams &p, int tid, int k0, bool tile_zlo, bool tile_zhi, int &ty, int &tx) {
    if (!(tile_zlo || tile_zhi)) { ty = tid >> 5; tx = (tid & 31) * 2; return true; }
    const int wcols = min(TX, p.n3 - k0);                                   // columns of the tile inside the grid
    const int nlo = tile_zlo ? min(p.P - k0, wcols) : 0;                    // PML columns [0, nlo)
    const int hi0 = tile_zhi ? max(p.n3 - p.P - k0, nlo) : wcols;           // PML columns [hi0, wcols)
    const int pend = (wcols + 1) >> 1;                                      // pairs of the tile inside the grid
    const int pl = min((nlo + 1) >> 1, pend);                               // pairs [0, pl) hold low-side PML columns
    const int ph0 = max(hi0 >> 1, pl);                                      // pairs [ph0, pend) hold high-side PML columns
    const int npp = pl + (pend - ph0), nip = ph0 - pl;
    int pair;
    if (tid < ROWS * npp) { ty = tid / npp; const int c = tid - ty * npp; pair = c < pl ? c : ph0 + (c - pl); }
    else if (tid - ROWS * npp < ROWS * nip) { const int t2 = tid - ROWS * npp; ty = t2 / nip; pair = pl + (t2 - ty * nip); }
    else { ty = 0; tx = 0; return false; }                                  // pairs beyond the grid: nothing to do
    tx = 2 * pair;
    return true;
}

// =========================================================================================
// stress half-step, two cells per thread
// =========================================================================================
template <typename LT, int ACC>
__global__ void __launch_bounds__(NTB2, 1) stress_tma2(const __grid_constant__ StressMaps tm, const DevParams p, const ChunkPlan plan) {
    constexpr bool SMC = sizeof(LT) == 1;
    constexpr int LW = LabBox<LT>::W;
    extern __shared__ __align__(1024) unsigned char sm[];   // TMA destinations need 128-byte alignment
    MatCoef *sC = reinterpret_cast<MatCoef *>(sm + OFF_COEF);
    AxisCoef *sJ = reinterpret_cast<AxisCoef *>(sm + OFF_AXJ);
    AxisCoef *sK = reinterpret_cast<AxisCoef *>(sm + OFF_AXK);
    unsigned char *sF = sm + OFF_FLAGS;
    const uint32_t sm32 = smem_u32(sm);
    const uint32_t fullH = sm32 + OFF_BAR, emptyH = fullH + MAX_NSH * 8, fullP = emptyH + MAX_NSH * 8, emptyP = fullP + MAX_NSP * 8;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k0 = blockIdx.x * TX, j0 = blockIdx.y * TY;
    const int ic0 = plan.start[blockIdx.z], ic1 = plan.end[blockIdx.z];
    const int np = ic1 - ic0;                 // planes of this CTA
    const unsigned long long dbg_t0 = (p.dbg && tid == 0) ? globaltimer_ns() : 0ull;
    const int ipl0 = ic0 - p.i0 + 2;          // local plane of ic0
    const bool tile_jd = (int)blockIdx.y < p.nylo || (int)blockIdx.y >= p.tjhi0;
    const bool tile_zlo = k0 < p.P, tile_zhi = k0 + TX > p.n3 - p.P;
    const int yt = ((int)blockIdx.y < p.nylo ? (int)blockIdx.y : (int)blockIdx.y - p.tjhi0 + p.nylo) * TY;
    int fl = 0;
    if (tid < np + 2) {
        const int ipl = ipl0 + tid;
        fl = ipl < p.nloc ? p.flags[((long long)ipl * p.ntj + blockIdx.y) * p.ntk + blockIdx.x] : 0;
        sF[tid] = (unsigned char)fl;
    }
    const bool any_solid = __syncthreads_or(tid < np && (fl & TF_SOLID)) != 0;
    // point-stage layout: identical to stress_tma (the producers are the same)
    const int zcomp = p.zbw * TY;
    const int zshear = align128(3 * zcomp * 4);
    const int zreg = zshear + (any_solid ? align128(2 * zcomp * 4) : 0);
    const int yoff = (any_solid ? PB_PARTS : PB_FLUID) * PBOX, zlo_off = yoff + (tile_jd ? (any_solid ? 5 : 3) * PBOX : 0);
    const int zhi_off = zlo_off + (tile_zlo ? zreg : 0);
    const int pstage = zhi_off + (tile_zhi ? zreg : 0);
    int nsp, nsh;
    ring_depths(pstage, ST_HSTAGE, nsp, nsh);
    const int offP = OFF_RINGS, offH = OFF_RINGS + nsp * pstage;

    if (SMC) for (int t = tid; t < p.nmat * (int)(sizeof(MatCoef) / 4); t += NTB2) reinterpret_cast<float *>(sC)[t] = reinterpret_cast<const float *>(p.coef)[t];
    if (tid < TY * 8) { const int r = tid >> 3, e = tid & 7; reinterpret_cast<float *>(sJ)[tid] = reinterpret_cast<const float *>(p.axJ + min(j0 + r, p.n2 - 1))[e]; }
    for (int t = tid; t < TX * 8; t += NTB2) { const int r = t >> 3, e = t & 7; reinterpret_cast<float *>(sK)[t] = reinterpret_cast<const float *>(p.axK + min(k0 + r, p.n3 - 1))[e]; }
    const bool first_hs = (unsigned)(p.seq & 0xffffffffu) <= 1u;
    const bool near_lo = ic0 < p.i0 + 2 && p.peerV[0] != nullptr, near_hi = ic1 > p.i1 - 2 && p.peerV[1] != nullptr;
    const bool has_peer = (ic0 < p.i0 + 2 || ic1 > p.i1 - 2) && (p.peerV[0] != nullptr || p.peerV[1] != nullptr);
    if (tid == 0) {
        if (!first_hs && near_lo) peer_wait(p, 0);
        if (!first_hs && near_hi) peer_wait(p, 1);
        for (int s = 0; s < nsh; s++) { mbar_init(fullH + s * 8, 1); mbar_init(emptyH + s * 8, NCW2); }
        for (int s = 0; s < nsp; s++) { mbar_init(fullP + s * 8, 1); mbar_init(emptyP + s * 8, NCW2); }
        fence_barrier_init();
    }
    __syncthreads();

    // =============================== producer warps (as in stress_tma) ===============================
    if (warp == NCW2) {
        if (lane != 0) return;
        RingPos rh(nsh, 0, 1);
        for (int r = 0; r < np + 2; r++) {
            const int slot = rh.slot;
            mbar_wait(emptyH + slot * 8, rh.par);
            const uint32_t st = sm32 + offH + slot * ST_HSTAGE;
            const uint32_t bar = fullH + slot * 8;
            mbar_expect_tx(bar, 3 * HBOX + LW * LH * (int)sizeof(LT));
            const int ipl = ipl0 + r;
            tma_load_4d(st, &tm.v3, bar, k0 - HK, j0 - HALO, ipl, 0);
            tma_load_3d(st + ST_LOFF, &tm.lab, bar, k0, j0, ipl);
            rh.advance();
        }
        return;
    }
    if (warp == NCW2 + 1) {
        if (lane != 0) return;
        RingPos rp(nsp, 0, 1);
        for (int r = 0; r < np; r++) {
            const int slot = rp.slot;
            mbar_wait(emptyP + slot * 8, rp.par);
            const uint32_t st = sm32 + offP + slot * pstage;
            const uint32_t bar = fullP + slot * 8;
            const unsigned f = sF[r];
            const int i = ic0 + r, ipl = ipl0 + r, io = i - p.i0;
            const bool xd = in_pml1(i, p.n1, p.P);
            const bool fint = f & TF_INT, fatt = f & TF_ATT, fsol = f & TF_SOLID;
            const int npart = fsol ? 5 : 3;
            const bool acc = ACC == 1 && fint;
            const int nbox = 3 + (fsol ? 3 : 0) + (xd ? npart : (fint ? 1 : 0) + (fatt ? 3 : 0) + (fsol && fint ? 3 : 0))
                           + (tile_jd ? npart : 0) + (acc ? 1 : 0);
            mbar_expect_tx(bar, nbox * PBOX + ((tile_zlo ? npart : 0) + (tile_zhi ? npart : 0)) * zcomp * 4);
            tma_load_4d(st + PB_SXX * PBOX, &tm.s3, bar, k0, j0, ipl, 0);
            if (fsol) tma_load_4d(st + PB_SXY * PBOX, &tm.s3, bar, k0, j0, ipl, 3);
            if (xd) {
                const int ipx = i < p.P ? io : p.nxlo + (i - p.xhi_begin);
                tma_load_4d(st + PB_RXX * PBOX, &tm.xp3, bar, k0, j0, ipx, 0);
                if (fsol) tma_load_4d(st + (PB_RXX + 3) * PBOX, &tm.xp2, bar, k0, j0, ipx, 3);
            } else {
                if (fint) tma_load_3d(st + PB_PR * PBOX, &tm.pr, bar, k0, j0, ipl);
                if (fatt) tma_load_4d(st + PB_RXX * PBOX, &tm.r3, bar, k0, j0, ipl, 0);
                if (fsol && fint) tma_load_4d(st + PB_RXY * PBOX, &tm.r3, bar, k0, j0, ipl, 3);
            }
            if (tile_jd) {
                tma_load_4d(st + yoff, &tm.yp3, bar, k0, yt, io, 0);
                if (fsol) tma_load_4d(st + yoff + 3 * PBOX, &tm.yp2, bar, k0, yt, io, 3);
            }
            if (tile_zlo) {
                tma_load_4d(st + zlo_off, &tm.zp3, bar, 0, j0, io, 0);
                if (fsol) tma_load_4d(st + zlo_off + zshear, &tm.zp2, bar, 0, j0, io, 3);
            }
            if (tile_zhi) {
                tma_load_4d(st + zhi_off, &tm.zp3, bar, p.zbw, j0, io, 0);
                if (fsol) tma_load_4d(st + zhi_off + zshear, &tm.zp2, bar, p.zbw, j0, io, 3);
            }
            if (acc) tma_load_3d(st + PB_ACC * PBOX, &tm.acc, bar, k0, j0, io);
            rp.advance();
        }
        return;
    }

    // =============================== consumer warps ===============================
    int tx, ty;
    const bool mapped = map_pair(p, tid, k0, tile_zlo, tile_zhi, ty, tx);
    const int k = k0 + tx, j = j0 + ty;                 // cell a = (j, k), cell b = (j, k + 1)
    const bool jd = in_pml1(j, p.n2, p.P);
    bool active[2], kd[2], upd_pml[2];
    float ckb_a[2], ckb_b[2], ckf_a[2], ckf_b[2];
#pragma unroll
    for (int c = 0; c < 2; c++) {
        active[c] = mapped && k + c < p.n3 && j < p.n2;
        kd[c] = in_pml1(k + c, p.n3, p.P);
        upd_pml[c] = active[c] && j < p.n2 - 1 && k + c < p.n3 - 1;
        ckb_a[c] = sK[tx + c].cab; ckb_b[c] = sK[tx + c].cbb; ckf_a[c] = sK[tx + c].caf; ckf_b[c] = sK[tx + c].cbf;
    }
    const float cjb_a = sJ[ty].cab, cjb_b = sJ[ty].cbb, cjf_a = sJ[ty].caf, cjf_b = sJ[ty].cbf;
    unsigned s1 = (unsigned)p.plane;   // element indices fit 32 bits (checked at create)
    keep(s1);

    // ---- register queues along i (state before the shift of plane ic0), one float2 per queue entry
    const float *__restrict__ Vx = p.V[0], *__restrict__ Vy = p.V[1], *__restrict__ Vz = p.V[2];
    const unsigned col = (unsigned)min(j, p.n2 - 1) * p.pitch + min(k, p.pitch - 2);
    unsigned q = (unsigned)ipl0 * s1 + col;
    float2 vx_m2, vx_m1 = ld2(Vx + q - 2 * s1), vx_0 = ld2(Vx + q - s1), vx_p1;
    float2 vy_m1, vy_0 = ld2(Vy + q - s1), vy_p1, vy_p2;
    float2 vz_m1, vz_0 = ld2(Vz + q - s1), vz_p1, vz_p2;
    const int sc = (ty + HALO) * SW + tx + HK;     // cell a in a halo box
    const int lc = ty * LW + tx;                     // ... in a label box
    const int pc = ty * TX + tx;                     // ... in a point box
    const float dt = p.dt;
    const unsigned MSK = LabelTraits<LT>::MASK;
    unsigned qy_stride = (unsigned)p.nyrows * p.pitch, qz_stride = (unsigned)p.n2 * p.zpw;
    int pushsel = has_peer ? ((ic0 < p.i0 + 2 && p.peerS[0] ? 1 : 0) | (ic1 > p.i1 - 2 && p.peerS[1] ? 2 : 0)) : 0;
    keep(pushsel);
    int nplanes = np, lane0 = lane == 0;
    keep(nplanes); keep(lane0);
    unsigned qy = ((unsigned)(ic0 - p.i0) * p.nyrows + yt + ty) * p.pitch + k;
    unsigned qz[2];
    int zsrc[2];
#pragma unroll
    for (int c = 0; c < 2; c++) {
        const int kc = k + c;
        const int kz = kc < p.P ? kc : kc - (p.n3 - p.P);                      // column inside the Z part box of the cell's side
        qz[c] = ((unsigned)(ic0 - p.i0) * p.n2 + min(j, p.n2 - 1)) * p.zpw + (kc < p.P ? kz : p.zbw + kz);
        zsrc[c] = ((kc < p.P ? zlo_off : zhi_off) >> 2) + ty * p.zbw + (kd[c] ? kz : 0);   // float offset inside a point stage
    }
    const bool jkd_any = jd || kd[0] || kd[1];

    const char *hsc = reinterpret_cast<const char *>(sm + offH) + sc * 4;
    const char *lsc = reinterpret_cast<const char *>(sm + offH + ST_LOFF) + lc * sizeof(LT);
    const char *psc = reinterpret_cast<const char *>(sm + offP) + pc * 4;
    const float *pzb = reinterpret_cast<const float *>(sm + offP);
    auto hbox = [&](int off, int c) { return reinterpret_cast<const float *>(hsc + off + c * HBOX_STRIDE); };
    auto lbox = [&](int off) { return reinterpret_cast<const LT *>(lsc + off); };
    int ho = 0, ho1 = ST_HSTAGE, ho2 = 2 * ST_HSTAGE, po = 0;
    uint32_t hb0 = fullH, hb1 = fullH + 8, hb2 = fullH + 16, pbar = fullP;
    unsigned hpar = 0, ppar = 0;
    const int hend = nsh * ST_HSTAGE, pend = nsp * pstage;

    mbar_wait(fullH, 0);
    vx_p1 = ld2(hbox(0, 0)); vy_p1 = ld2(hbox(0, 1)); vz_p1 = ld2(hbox(0, 2));
    mbar_wait(fullH + 8, 0);
    vy_p2 = ld2(hbox(ST_HSTAGE, 1)); vz_p2 = ld2(hbox(ST_HSTAGE, 2));

    for (int it = 0; it < nplanes; it++, q += s1, qy += qy_stride, qz[0] += qz_stride, qz[1] += qz_stride) {
        const int i = ic0 + it;
        const unsigned f = sF[it];
        mbar_wait(hb2, hpar);
        vx_m2 = vx_m1; vx_m1 = vx_0; vx_0 = vx_p1; vx_p1 = ld2(hbox(ho1, 0));
        vy_m1 = vy_0; vy_0 = vy_p1; vy_p1 = vy_p2; vy_p2 = ld2(hbox(ho2, 1));
        vz_m1 = vz_0; vz_0 = vz_p1; vz_p1 = vz_p2; vz_p2 = ld2(hbox(ho2, 2));
        mbar_wait(pbar, ppar);
        const bool xd = (f & TF_XD) != 0;
        const bool ilast = (f & TF_ILAST) != 0;
        // which of the two cells this thread updates on this plane
        bool cpml[2], upd[2];
#pragma unroll
        for (int c = 0; c < 2; c++) {
            cpml[c] = xd || jd || kd[c];
            upd[c] = cpml[c] ? (upd_pml[c] && !ilast) : active[c];
        }
        auto pair_update = [&](auto solid_tag) {
            constexpr bool SOL = decltype(solid_tag)::value;
            const float *bx = hbox(ho, 0), *by = hbox(ho, 1), *bz = hbox(ho, 2);
            const LT *l0p = lbox(ho), *l1p = lbox(ho1);
            float *pb = const_cast<float *>(reinterpret_cast<const float *>(psc + po));
            float cib_a = BB_CA, cib_b = BB_CB, cif_a = BB_CA, cif_b = BB_CB;
            if (f & TF_IEDGE) { const AxisCoef ci = load_axis(p.axI, i); cib_a = ci.cab; cib_b = ci.cbb; cif_a = ci.caf; cif_b = ci.cbf; }
            // ---------------- staggered differences of both cells; the k-stencils share their taps
            float D[2][9];
            {
                const float2 y0 = ld2(by), ym1 = ld2(by - SW), yp1 = ld2(by + SW), ym2 = ld2(by - 2 * SW);
                const float2 zm = ld2(bz - 2), z0 = ld2(bz);
                const float z2 = bz[2];
                D[0][0] = D4C(cib_a, cib_b, vx_0.x, vx_m1.x, vx_p1.x, vx_m2.x);
                D[1][0] = D4C(cib_a, cib_b, vx_0.y, vx_m1.y, vx_p1.y, vx_m2.y);
                D[0][1] = D4C(cjb_a, cjb_b, y0.x, ym1.x, yp1.x, ym2.x);
                D[1][1] = D4C(cjb_a, cjb_b, y0.y, ym1.y, yp1.y, ym2.y);
                D[0][2] = D4C(ckb_a[0], ckb_b[0], z0.x, zm.y, z0.y, zm.x);
                D[1][2] = D4C(ckb_a[1], ckb_b[1], z0.y, z0.x, z2, zm.y);
                if constexpr (SOL) {
                    const float2 x0 = ld2(bx), xp1 = ld2(bx + SW), xp2 = ld2(bx + 2 * SW), xm1 = ld2(bx - SW), x2 = ld2(bx + 2);
                    const float xk1 = bx[-1];
                    const float2 zp1 = ld2(bz + SW), zp2 = ld2(bz + 2 * SW), zm1 = ld2(bz - SW);
                    const float2 y2 = ld2(by + 2);
                    const float yk1 = by[-1];
                    D[0][3] = D4C(cif_a, cif_b, vy_p1.x, vy_0.x, vy_p2.x, vy_m1.x);
                    D[1][3] = D4C(cif_a, cif_b, vy_p1.y, vy_0.y, vy_p2.y, vy_m1.y);
                    D[0][4] = D4C(cjf_a, cjf_b, xp1.x, x0.x, xp2.x, xm1.x);
                    D[1][4] = D4C(cjf_a, cjf_b, xp1.y, x0.y, xp2.y, xm1.y);
                    D[0][5] = D4C(cif_a, cif_b, vz_p1.x, vz_0.x, vz_p2.x, vz_m1.x);
                    D[1][5] = D4C(cif_a, cif_b, vz_p1.y, vz_0.y, vz_p2.y, vz_m1.y);
                    D[0][6] = D4C(ckf_a[0], ckf_b[0], x0.y, x0.x, x2.x, xk1);
                    D[1][6] = D4C(ckf_a[1], ckf_b[1], x2.x, x0.y, x2.y, x0.x);
                    D[0][7] = D4C(cjf_a, cjf_b, zp1.x, z0.x, zp2.x, zm1.x);
                    D[1][7] = D4C(cjf_a, cjf_b, zp1.y, z0.y, zp2.y, zm1.y);
                    D[0][8] = D4C(ckf_a[0], ckf_b[0], y0.y, y0.x, y2.x, yk1);
                    D[1][8] = D4C(ckf_a[1], ckf_b[1], y2.x, y0.y, y2.y, y0.x);
                } else {
#pragma unroll
                    for (int c = 0; c < 2; c++) { D[c][3] = D[c][4] = D[c][5] = D[c][6] = D[c][7] = D[c][8] = 0.f; }
                }
            }
            // ---------------- the fields of the pair as staged (old values), updated in registers, written back as pairs
            float s[2][6], r[2][6], pr[2], ac[2];
            {
                const float2 a = ld2(pb + PB_SXX * BOXF), b = ld2(pb + PB_SYY * BOXF), c = ld2(pb + PB_SZZ * BOXF);
                s[0][0] = a.x; s[1][0] = a.y; s[0][1] = b.x; s[1][1] = b.y; s[0][2] = c.x; s[1][2] = c.y;
                if constexpr (SOL) {
                    const float2 d = ld2(pb + PB_SXY * BOXF), e = ld2(pb + PB_SXZ * BOXF), g = ld2(pb + PB_SYZ * BOXF);
                    s[0][3] = d.x; s[1][3] = d.y; s[0][4] = e.x; s[1][4] = e.y; s[0][5] = g.x; s[1][5] = g.y;
                } else {
#pragma unroll
                    for (int c2 = 0; c2 < 2; c2++) { s[c2][3] = s[c2][4] = s[c2][5] = 0.f; }
                }
            }
            const bool any_int = (upd[0] && !cpml[0]) || (upd[1] && !cpml[1]);      // the pair holds an interior cell of this plane
            bool wr_r = false, wr_sh[3] = { false, false, false }, wr_rsh[3] = { false, false, false };
            bool wr_shear_all = false;
            if (any_int) {
                const float2 a = ld2(pb + PB_PR * BOXF);
                pr[0] = a.x; pr[1] = a.y;
                if (ACC == 1) { const float2 b = ld2(pb + PB_ACC * BOXF); ac[0] = b.x; ac[1] = b.y; }
            }
#pragma unroll
            for (int c = 0; c < 2; c++) {
                if (!upd[c]) continue;
                const unsigned l0 = l0p[c];
                const bool refl = (l0 & LabelTraits<LT>::REFL) != 0;
                MatCoef mc;
                if (SMC) mc = sC[l0 & MSK]; else mc = load_coef_global(p.coef, l0 & MSK);
                float rigxy = 0.f, rigxz = 0.f, rigyz = 0.f, texy = 0.f, texz = 0.f, teyz = 0.f;
                if constexpr (SOL) {
                    const unsigned mi = l1p[c] & MSK, mj = l0p[LW + c] & MSK, mk = l0p[1 + c] & MSK;
                    const unsigned mij = l1p[LW + c] & MSK, mik = l1p[1 + c] & MSK, mjk = l0p[LW + 1 + c] & MSK;
                    float igi, igj, igk, igij, igik, igjk, ti, tj, tk, tij, tik, tjk;
                    if (SMC) {
                        igi = sC[mi].invG; igj = sC[mj].invG; igk = sC[mk].invG; igij = sC[mij].invG; igik = sC[mik].invG; igjk = sC[mjk].invG;
                        ti = sC[mi].tauS; tj = sC[mj].tauS; tk = sC[mk].tauS; tij = sC[mij].tauS; tik = sC[mik].tauS; tjk = sC[mjk].tauS;
                    } else {
                        igi = __ldg(&p.coef[mi].invG); igj = __ldg(&p.coef[mj].invG); igk = __ldg(&p.coef[mk].invG);
                        igij = __ldg(&p.coef[mij].invG); igik = __ldg(&p.coef[mik].invG); igjk = __ldg(&p.coef[mjk].invG);
                        ti = __ldg(&p.coef[mi].tauS); tj = __ldg(&p.coef[mj].tauS); tk = __ldg(&p.coef[mk].tauS);
                        tij = __ldg(&p.coef[mij].tauS); tik = __ldg(&p.coef[mik].tauS); tjk = __ldg(&p.coef[mjk].tauS);
                    }
                    rigxy = rigidity4(mc.invG, igi, igj, igij);
                    rigxz = rigidity4(mc.invG, igi, igk, igik);
                    rigyz = rigidity4(mc.invG, igj, igk, igjk);
                    texy = 0.25f * (mc.tauS + ti + tj + tij);
                    texz = 0.25f * (mc.tauS + ti + tk + tik);
                    teyz = 0.25f * (mc.tauS + tj + tk + tjk);
                }
                if (cpml[c]) {
                    // ---------------- PML shell: damped split parts (old values staged by TMA), per cell as in stress_tma
                    PmlCell pcell;
                    pcell.xd = xd; pcell.jd = jd; pcell.kd = kd[c];
                    const int ipx = i < p.P ? i - p.i0 : p.nxlo + (i - p.xhi_begin);
                    pcell.qx = (unsigned)ipx * s1 + col + c; pcell.qy = qy + c; pcell.qz = qz[c];
                    pcell.cI = p.axI + i; pcell.cJ = sJ + ty; pcell.cK = sK + tx + c;
                    const float *oz = pzb + (po >> 2) + zsrc[c];
                    stress_pml<true>(p, pcell, mc.M, mc.L, rigxy, rigxz, rigyz, D[c], s[c], pb + PB_RXX * BOXF + c, pb + (yoff >> 2) + c, oz,
                                     oz + (zshear >> 2), BOXF, zcomp);
                    if (refl) { s[c][0] = s[c][1] = s[c][2] = s[c][3] = s[c][4] = s[c][5] = 0.f; }
                    if (SOL) wr_shear_all = true;
                } else {
                    // ---------------- interior: viscoelastic update
                    const bool att = attenuates(mc);
                    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
                    if (att) { r0 = pb[PB_RXX * BOXF + c]; r1 = pb[PB_RYY * BOXF + c]; r2 = pb[PB_RZZ * BOXF + c]; }
                    stress_normal_interior(mc, dt, att, D[c][0], D[c][1], D[c][2], s[c][0], s[c][1], s[c][2], r0, r1, r2, pr[c]);
                    if (refl) { s[c][0] = s[c][1] = s[c][2] = 0.f; pr[c] = 0.f; }
                    if (att) {
                        // the partner's memory variables are written back as staged unless it updates them itself
                        if (!wr_r) {
                            const float2 a = ld2(pb + PB_RXX * BOXF), b = ld2(pb + PB_RYY * BOXF), e = ld2(pb + PB_RZZ * BOXF);
                            r[0][0] = a.x; r[1][0] = a.y; r[0][1] = b.x; r[1][1] = b.y; r[0][2] = e.x; r[1][2] = e.y;
                            wr_r = true;
                        }
                        r[c][0] = r0; r[c][1] = r1; r[c][2] = r2;
                    }
                    if constexpr (SOL) {
                        const float rig[3] = { rigxy, rigxz, rigyz }, te[3] = { texy, texz, teyz };
                        const float Ds[3] = { D[c][3] + D[c][4], D[c][5] + D[c][6], D[c][7] + D[c][8] };
#pragma unroll
                        for (int n = 0; n < 3; n++) {
                            if (rig[n] != 0.f) {
                                float rr = pb[(PB_RXY + n) * BOXF + c];
                                stress_shear_interior(mc, dt, rig[n], te[n], Ds[n], s[c][3 + n], rr);
                                if (refl) s[c][3 + n] = 0.f;
                                wr_sh[n] = true;
                                if (te[n] != 0.f) {
                                    if (!wr_rsh[n]) { const float2 a = ld2(pb + (PB_RXY + n) * BOXF); r[0][3 + n] = a.x; r[1][3 + n] = a.y; wr_rsh[n] = true; }
                                    r[c][3 + n] = rr;
                                }
                            }
                        }
                    }
                    if (ACC == 1) {
                        const float v = -mc.K * pr[c];
                        ac[c] += v * v;
                    } else if (ACC == 2) {
                        const unsigned qa = q + c - 2 * s1;
#pragma unroll
                        for (int n = 0; n < 6; n++) accumulate(p, BB_MAP_SXX + n, qa, s[c][n], false);
                        accumulate(p, BB_MAP_PRESSURE, qa, -mc.K * pr[c], false);
                    }
                }
            }
            // ---------------- write the pair back: one 64-bit store per field (cells that did not change a field keep its staged value)
            st2(p.S[0] + q, s[0][0], s[1][0]); st2(p.S[1] + q, s[0][1], s[1][1]); st2(p.S[2] + q, s[0][2], s[1][2]);
            if (any_int) {
                st2(p.Pr + q, pr[0], pr[1]);
                if (ACC == 1) st2(p.acc_rms + (q - 2 * s1), ac[0], ac[1]);
            }
            if (wr_r) { st2(p.R[0] + q, r[0][0], r[1][0]); st2(p.R[1] + q, r[0][1], r[1][1]); st2(p.R[2] + q, r[0][2], r[1][2]); }
            if constexpr (SOL) {
#pragma unroll
                for (int n = 0; n < 3; n++) {
                    if (wr_shear_all || wr_sh[n]) st2(p.S[3 + n] + q, s[0][3 + n], s[1][3 + n]);
                    if (wr_rsh[n]) st2(p.R[3 + n] + q, r[0][3 + n], r[1][3 + n]);
                }
            }
            // ---------------- boundary planes also go to the slab neighbour (the stresses its particle update differentiates along i)
            if (pushsel) {
                if (p.bsrc_map) {      // stress sources of the boundary planes are injected here (the source kernel skips them)
                    const int bp = boundary_plane(p, pushsel, i);
                    bool any = false;
#pragma unroll
                    for (int c = 0; c < 2; c++) {
                        float val, ox, oy, oz;
                        if (bp >= 0 && active[c] && boundary_source(p, bp, col + c, val, ox, oy, oz)) {
                            const float w = val * ox;
                            if (p.src_hard) { s[c][0] = w; s[c][1] = w; s[c][2] = w; } else { s[c][0] += w; s[c][1] += w; s[c][2] += w; }
                            any = true;
                        }
                    }
                    if (any) { st2(p.S[0] + q, s[0][0], s[1][0]); st2(p.S[1] + q, s[0][1], s[1][1]); st2(p.S[2] + q, s[0][2], s[1][2]); }
                }
                if ((pushsel & 1) && i < p.i0 + 2) {
                    float *b = p.peerS[0];
                    const unsigned qn = (p.peer_plane[0] + (unsigned)(i - p.i0)) * s1 + col;
                    st2(b + qn, s[0][0], s[1][0]);
                    if constexpr (SOL) { st2(b + 3 * p.peer_vol[0] + qn, s[0][3], s[1][3]); st2(b + 4 * p.peer_vol[0] + qn, s[0][4], s[1][4]); }
                }
                if ((pushsel & 2) && i >= p.i1 - 2) {
                    float *b = p.peerS[1];
                    const unsigned qn = (p.peer_plane[1] + (unsigned)(i - (p.i1 - 2))) * s1 + col;
                    st2(b + qn, s[0][0], s[1][0]);
                    if constexpr (SOL) { st2(b + 3 * p.peer_vol[1] + qn, s[0][3], s[1][3]); st2(b + 4 * p.peer_vol[1] + qn, s[0][4], s[1][4]); }
                }
            }
        };
        if (upd[0] || upd[1]) {
            if (f & TF_SOLID) pair_update(std::true_type{}); else pair_update(std::false_type{});
        }
        // ---------------- a slab-boundary plane pair of this CTA is complete: count it in right away
        if (pushsel) {
            const bool last_lo = (pushsel & 1) && i == min(ic1, p.i0 + 2) - 1;
            const bool last_hi = (pushsel & 2) && i == ic1 - 1;
            if ((last_lo || last_hi) && p.publish) {
                consumer_bar2();
                if (tid == 0) {
                    __threadfence_system();
                    const unsigned expected = 2u * gridDim.x * gridDim.y;
                    if (last_lo) peer_publish(p, 0, (unsigned)(min(ic1, p.i0 + 2) - ic0), expected);
                    if (last_hi) peer_publish(p, 1, (unsigned)(ic1 - max(ic0, p.i1 - 2)), expected);
                }
            }
        }
        // ---------------- this warp is done with the slots of plane i
        __syncwarp();
        if (lane0) { mbar_arrive(hb0 + MAX_NSH * 8); mbar_arrive(pbar + MAX_NSP * 8); }
        ho = ho1; hb0 = hb1; ho1 = ho2; hb1 = hb2;
        ho2 += ST_HSTAGE; hb2 += 8;
        if (ho2 == hend) { ho2 = 0; hb2 = fullH; hpar ^= 1u; }
        po += pstage; pbar += 8;
        if (po == pend) { po = 0; pbar = fullP; ppar ^= 1u; }
    }
    (void)jkd_any;
    if (p.dbg && tid == 0) {
        unsigned long long *d = p.dbg + 4ull * ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x);
        d[0] = dbg_t0; d[1] = globaltimer_ns(); d[2] = ((unsigned long long)blockIdx.z << 40) | ((unsigned long long)blockIdx.y << 20) | blockIdx.x; d[3] = (unsigned long long)np;
    }
}

// =========================================================================================
// particle half-step, two cells per thread, tiles of ROWS x 64 cells (ROWS = 8: the tiling of every other kernel;
// ROWS = 16: the production particle kernel -- its stages are small (three velocity boxes, two to five stress halo
// boxes), so twice the rows fit the same shared memory, the CTA keeps 16 consumer warps with a cell pair per thread,
// and the (ROWS+4)/ROWS halo over-read drops from 1.5 to 1.25)
// =========================================================================================
template <int ROWS> struct PT {
    // ROWS = 16: no producer warps -- lane 0 of consumer warps 0 and 1 issues the TMA loads of the two rings between two
    // of its own planes.  16 warps are 4 per scheduler and leave 128 registers per thread; with 18 warps one scheduler
    // holds 5 and ptxas has to fit a cell pair into 96 registers (it spilled ~30 values inside the plane loop:
    // profiles/r2_kernel_experiments.txt).
    static constexpr bool FOLD = ROWS >= 16;
    static constexpr int NCW = ROWS, NT = ROWS * 32, NTB = NT + (FOLD ? 0 : 64);
    static constexpr int SHH = ROWS + 2 * HALO, HBOXB = SW * SHH * 4, PBOXB = TX * ROWS * 4, BOXFL = TX * ROWS;
    static constexpr int LHH = ROWS + 1, LBOXB = ((TX + 8) * 2 * LHH + 127) & ~127;
    static constexpr int XOFF = align128(2 * HBOXB), LOFF = XOFF + PBOXB, S3OFF = align128(LOFF + LBOXB), HSTAGE = S3OFF + align128(3 * HBOXB);
    static constexpr int AXK = OFF_AXJ + ROWS * (int)sizeof(AxisCoef), FLAGS = AXK + TX * (int)sizeof(AxisCoef), BAR = FLAGS + ((MAXCHUNK + 8 + 15) / 16) * 16;
    static_assert(BAR + 2 * (MAX_NSH + MAX_NSP) * 8 <= OFF_RINGS, "tables overflow their 9 KB");
};
template <typename LT, int ACC, int ROWS>
__global__ void __launch_bounds__(PT<ROWS>::NTB, 1) particle_tma2(const __grid_constant__ ParticleMaps tm, const DevParams p, const ChunkPlan plan) {
    constexpr bool SMC = sizeof(LT) == 1;
    constexpr int LW = LabBox<LT>::W;
    extern __shared__ __align__(1024) unsigned char sm[];
    float *sB = reinterpret_cast<float *>(sm + OFF_COEF);
    AxisCoef *sJ = reinterpret_cast<AxisCoef *>(sm + OFF_AXJ);
    using G = PT<ROWS>;
    constexpr int HBOXR = G::HBOXB, PBOXR = G::PBOXB, BOXFR = G::BOXFL, NCWR = G::NCW, NTBR = G::NTB;
    AxisCoef *sK = reinterpret_cast<AxisCoef *>(sm + G::AXK);
    unsigned char *sF = sm + G::FLAGS;
    const uint32_t sm32 = smem_u32(sm);
    const uint32_t fullH = sm32 + G::BAR, emptyH = fullH + MAX_NSH * 8, fullP = emptyH + MAX_NSH * 8, emptyP = fullP + MAX_NSP * 8;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k0 = blockIdx.x * TX, j0 = blockIdx.y * ROWS;
    const int ic0 = plan.start[blockIdx.z], ic1 = plan.end[blockIdx.z];
    const int np = ic1 - ic0;
    const unsigned long long dbg_t0 = (p.dbg && tid == 0) ? globaltimer_ns() : 0ull;
    const int ipl0 = ic0 - p.i0 + 2;
    // a tile of ROWS rows is NSUB tiles of the granularity (TY rows) at which the traffic flags and the Y parts are kept:
    // its flags are the OR of theirs; its Y-part box starts at the storage row of its first sub-tile that is stored
    // (stored sub-tiles of one tile are consecutive in storage; rows of a sub-tile that is not stored hold no j-PML cell)
    constexpr int NSUB = ROWS / TY;
    static_assert(ROWS % TY == 0, "tile height must be a multiple of the flag granularity");
    const int t0 = (int)blockIdx.y * NSUB;
    bool tile_jd = false;
    int yt = 0;
#pragma unroll
    for (int u = NSUB - 1; u >= 0; u--) {
        const int t = t0 + u;
        if (t < p.ntj && (t < p.nylo || t >= p.tjhi0)) { tile_jd = true; yt = ((t < p.nylo ? t : t - p.tjhi0 + p.nylo) - u) * TY; }
    }
    const bool tile_zlo = k0 < p.P, tile_zhi = k0 + TX > p.n3 - p.P;
    int fl = 0;
    if (tid < np + 2) {
        const int ipl = ipl0 + tid;
        if (ipl < p.nloc) {
#pragma unroll
            for (int u = 0; u < NSUB; u++)
                if (t0 + u < p.ntj) fl |= p.flags[((long long)ipl * p.ntj + t0 + u) * p.ntk + blockIdx.x];
        }
        sF[tid] = (unsigned char)fl;
    }
    const bool any_shear = __syncthreads_or(fl & TF_SHEAR) != 0;
    const bool any_xd = ic0 < p.P || ic1 > p.n1 - p.P;
    const int hstage = any_shear ? G::HSTAGE : G::S3OFF;
    const int zcomp = p.zbw * ROWS;
    const int zreg = align128(3 * zcomp * 4);
    const int yoff = (any_xd ? QB_PARTS : QB_X) * PBOXR, zlo_off = yoff + (tile_jd ? 3 * PBOXR : 0), zhi_off = zlo_off + (tile_zlo ? zreg : 0);
    const int pstage = zhi_off + (tile_zhi ? zreg : 0);
    int nsp, nsh;
    ring_depths(pstage, hstage, nsp, nsh);
    const int offP = OFF_RINGS, offH = OFF_RINGS + nsp * pstage;

    if (SMC) for (int t = tid; t < p.nmat; t += NTBR) sB[t] = p.coef[t].B;
    if (tid < ROWS * 8) { const int r = tid >> 3, e = tid & 7; reinterpret_cast<float *>(sJ)[tid] = reinterpret_cast<const float *>(p.axJ + min(j0 + r, p.n2 - 1))[e]; }
    for (int t = tid; t < TX * 8; t += NTBR) { const int r = t >> 3, e = t & 7; reinterpret_cast<float *>(sK)[t] = reinterpret_cast<const float *>(p.axK + min(k0 + r, p.n3 - 1))[e]; }
    const bool first_hs = (unsigned)(p.seq & 0xffffffffu) <= 1u;
    const bool near_lo = ic0 < p.i0 + 2 && p.peerS[0] != nullptr, near_hi = ic1 > p.i1 - 2 && p.peerS[1] != nullptr;
    const bool has_peer = (ic0 < p.i0 + 2 || ic1 > p.i1 - 2) && (p.peerV[0] != nullptr || p.peerV[1] != nullptr);
    if (tid == 0) {
        if (!first_hs && near_lo) peer_wait(p, 0);
        if (!first_hs && near_hi) peer_wait(p, 1);
        for (int s = 0; s < nsh; s++) { mbar_init(fullH + s * 8, 1); mbar_init(emptyH + s * 8, NCWR); }
        for (int s = 0; s < nsp; s++) { mbar_init(fullP + s * 8, 1); mbar_init(emptyP + s * 8, NCWR); }
        fence_barrier_init();
    }
    __syncthreads();

    // =============================== TMA loads of one plane into one slot of each ring ===============================
    auto load_halo = [&](int r, int slot, unsigned par) {      // stresses with halo + Sxx + labels of plane ic0 + r
        mbar_wait(emptyH + slot * 8, par);
        const uint32_t st = sm32 + offH + slot * hstage;
        const uint32_t bar = fullH + slot * 8;
        const bool fsh = sF[r] & TF_SHEAR;
        mbar_expect_tx(bar, (fsh ? 5 : 2) * HBOXR + PBOXR + LW * G::LHH * (int)sizeof(LT));
        const int ipl = ipl0 + r;
        tma_load_4d(st, &tm.sh2, bar, k0 - HK, j0 - HALO, ipl, 1);
        if (fsh) tma_load_4d(st + G::S3OFF, &tm.sh3, bar, k0 - HK, j0 - HALO, ipl, 3);
        tma_load_4d(st + G::XOFF, &tm.sxx, bar, k0, j0, ipl, 0);
        tma_load_3d(st + G::LOFF, &tm.lab, bar, k0, j0, ipl);
    };
    auto load_point = [&](int r, int slot, unsigned par) {     // V and its damped parts of plane ic0 + r
        mbar_wait(emptyP + slot * 8, par);
        const uint32_t st = sm32 + offP + slot * pstage;
        const uint32_t bar = fullP + slot * 8;
        const int i = ic0 + r, ipl = ipl0 + r, io = i - p.i0;
        const bool xd = in_pml1(i, p.n1, p.P);
        mbar_expect_tx(bar, (3 + (xd ? 3 : 0) + (tile_jd ? 3 : 0)) * PBOXR + ((tile_zlo ? 3 : 0) + (tile_zhi ? 3 : 0)) * zcomp * 4);
        tma_load_4d(st + QB_V * PBOXR, &tm.v3, bar, k0, j0, ipl, 0);
        if (xd) tma_load_4d(st + QB_X * PBOXR, &tm.xp3, bar, k0, j0, i < p.P ? io : p.nxlo + (i - p.xhi_begin), 5);
        if (tile_jd) tma_load_4d(st + yoff, &tm.yp3, bar, k0, yt, io, 5);
        if (tile_zlo) tma_load_4d(st + zlo_off, &tm.zp3, bar, 0, j0, io, 5);
        if (tile_zhi) tma_load_4d(st + zhi_off, &tm.zp3, bar, p.zbw, j0, io, 5);
    };
    RingPos rh(nsh, 0, 1), rp(nsp, 0, 1);      // producer positions (used by the issuing lanes only)
    if constexpr (!G::FOLD) {
        // ---- two producer warps, as in particle_tma
        if (warp == NCWR) {
            if (lane != 0) return;
            for (int r = 0; r < np + 2; r++) { load_halo(r, rh.slot, rh.par); rh.advance(); }
            return;
        }
        if (warp == NCWR + 1) {
            if (lane != 0) return;
            for (int r = 0; r < np; r++) { load_point(r, rp.slot, rp.par); rp.advance(); }
            return;
        }
    } else {
        // ---- folded: fill the rings now; afterwards every plane a warp finishes frees one slot, which its lane 0 refills
        // (it waits for the other warps to release that slot: they are at most a few planes apart)
        if (warp == 0 && lane == 0) for (int r = 0; r < min(nsh, np + 2); r++) { load_halo(r, rh.slot, rh.par); rh.advance(); }
        if (warp == 1 && lane == 0) for (int r = 0; r < min(nsp, np); r++) { load_point(r, rp.slot, rp.par); rp.advance(); }
    }

    // =============================== consumer warps ===============================
    int tx, ty;
    const bool mapped = map_pair<ROWS>(p, tid, k0, tile_zlo, tile_zhi, ty, tx);
    const int k = k0 + tx, j = j0 + ty;
    const bool jd = in_pml1(j, p.n2, p.P);
    bool active[2], kd[2], upd_pml[2];
    float ckb_a[2], ckb_b[2], ckf_a[2], ckf_b[2];
#pragma unroll
    for (int c = 0; c < 2; c++) {
        active[c] = mapped && k + c < p.n3 && j < p.n2;
        kd[c] = in_pml1(k + c, p.n3, p.P);
        upd_pml[c] = active[c] && j < p.n2 - 1 && k + c < p.n3 - 1;
        ckb_a[c] = sK[tx + c].cab; ckb_b[c] = sK[tx + c].cbb; ckf_a[c] = sK[tx + c].caf; ckf_b[c] = sK[tx + c].cbf;
    }
    const float cjb_a = sJ[ty].cab, cjb_b = sJ[ty].cbb, cjf_a = sJ[ty].caf, cjf_b = sJ[ty].cbf;
    unsigned s1 = (unsigned)p.plane;
    keep(s1);

    // queues: Sxx holds i-1..i+2 ; Sxy, Sxz hold i-2..i+1 (state before the shift of plane ic0)
    const float *__restrict__ Sxx = p.S[0], *__restrict__ Sxy = p.S[3], *__restrict__ Sxz = p.S[4];
    const unsigned col = (unsigned)min(j, p.n2 - 1) * p.pitch + min(k, p.pitch - 2);
    unsigned q = (unsigned)ipl0 * s1 + col;
    const float2 zero2 = make_float2(0.f, 0.f);
    float2 xx_m1, xx_0 = ld2(Sxx + q - s1), xx_p1, xx_p2;
    float2 xy_m2, xy_m1 = ld2(Sxy + q - 2 * s1), xy_0 = ld2(Sxy + q - s1), xy_p1 = zero2;
    float2 xz_m2, xz_m1 = ld2(Sxz + q - 2 * s1), xz_0 = ld2(Sxz + q - s1), xz_p1 = zero2;
    const int sc = (ty + HALO) * SW + tx + HK;
    const int lc = ty * LW + tx;
    const int pc = ty * TX + tx;
    const float dt = p.dt;
    const unsigned MSK = LabelTraits<LT>::MASK;
    unsigned qy_stride = (unsigned)p.nyrows * p.pitch, qz_stride = (unsigned)p.n2 * p.zpw;
    int pushsel = has_peer ? ((ic0 < p.i0 + 2 && p.peerV[0] ? 1 : 0) | (ic1 > p.i1 - 2 && p.peerV[1] ? 2 : 0)) : 0;
    keep(pushsel);
    int nplanes = np, lane0 = lane == 0;
    keep(nplanes); keep(lane0);
    const int tj8 = min(j, p.n2 - 1) / TY;                                 // this row's sub-tile and its storage row (meaningful for j-PML rows only)
    const int jp = (tj8 < p.nylo ? tj8 : tj8 - p.tjhi0 + p.nylo) * TY + (min(j, p.n2 - 1) - tj8 * TY);
    unsigned qy = ((unsigned)(ic0 - p.i0) * p.nyrows + (jd ? jp : 0)) * p.pitch + k;
    unsigned qz[2];
    int zsrc[2];
#pragma unroll
    for (int c = 0; c < 2; c++) {
        const int kc = k + c;
        const int kz = kc < p.P ? kc : kc - (p.n3 - p.P);
        qz[c] = ((unsigned)(ic0 - p.i0) * p.n2 + min(j, p.n2 - 1)) * p.zpw + (kc < p.P ? kz : p.zbw + kz);
        zsrc[c] = ((kc < p.P ? zlo_off : zhi_off) >> 2) + ty * p.zbw + (kd[c] ? kz : 0);
    }

    const char *hsc = reinterpret_cast<const char *>(sm + offH) + sc * 4;
    const char *xsc = reinterpret_cast<const char *>(sm + offH + G::XOFF) + pc * 4;
    const char *lsc = reinterpret_cast<const char *>(sm + offH + G::LOFF) + lc * sizeof(LT);
    const char *psc = reinterpret_cast<const char *>(sm + offP) + pc * 4;
    const float *pzb = reinterpret_cast<const float *>(sm + offP);
    auto hbox = [&](int off, int c) { return reinterpret_cast<const float *>(hsc + off + (c < 2 ? c * HBOXR : G::S3OFF + (c - 2) * HBOXR)); };
    auto xxbox = [&](int off) { return reinterpret_cast<const float *>(xsc + off); };
    auto lbox = [&](int off) { return reinterpret_cast<const LT *>(lsc + off); };
    int ho = 0, ho1 = hstage, ho2 = 2 * hstage, po = 0;
    uint32_t hb0 = fullH, hb1 = fullH + 8, hb2 = fullH + 16, pbar = fullP;
    unsigned hpar = 0, ppar = 0;
    const int hend = nsh * hstage, pend = nsp * pstage;

    mbar_wait(fullH, 0);
    xx_p1 = ld2(xxbox(0));
    if (sF[0] & TF_SHEAR) { xy_p1 = ld2(hbox(0, HB_SXY)); xz_p1 = ld2(hbox(0, HB_SXZ)); }
    mbar_wait(fullH + 8, 0);
    xx_p2 = ld2(xxbox(hstage));

    for (int it = 0; it < nplanes; it++, q += s1, qy += qy_stride, qz[0] += qz_stride, qz[1] += qz_stride) {
        const int i = ic0 + it;
        const unsigned f = sF[it];
        const bool fsh = f & TF_SHEAR;
        mbar_wait(hb2, hpar);
        xx_m1 = xx_0; xx_0 = xx_p1; xx_p1 = xx_p2; xx_p2 = ld2(xxbox(ho2));
        xy_m2 = xy_m1; xy_m1 = xy_0; xy_0 = xy_p1;
        xz_m2 = xz_m1; xz_m1 = xz_0; xz_0 = xz_p1;
        if (sF[it + 1] & TF_SHEAR) { xy_p1 = ld2(hbox(ho1, HB_SXY)); xz_p1 = ld2(hbox(ho1, HB_SXZ)); }
        else { xy_p1 = zero2; xz_p1 = zero2; }
        mbar_wait(pbar, ppar);
        const bool xd = (f & TF_XD) != 0;
        const bool ilast = (f & TF_ILAST) != 0;
        bool cpml[2], upd[2];
#pragma unroll
        for (int c = 0; c < 2; c++) {
            cpml[c] = xd || jd || kd[c];
            upd[c] = cpml[c] ? (upd_pml[c] && !ilast) : active[c];
        }
        auto pair_update = [&](auto shear_tag) {
            constexpr bool SHEAR = decltype(shear_tag)::value;
            const float *byy = hbox(ho, HB_SYY), *bzz = hbox(ho, HB_SZZ);
            const float *bxy = hbox(ho, HB_SXY), *bxz = hbox(ho, HB_SXZ), *byz = hbox(ho, HB_SYZ);
            const LT *l0p = lbox(ho), *l1p = lbox(ho1);
            float *pb = const_cast<float *>(reinterpret_cast<const float *>(psc + po));
            float cib_a = BB_CA, cib_b = BB_CB, cif_a = BB_CA, cif_b = BB_CB;
            if (f & TF_IEDGE) { const AxisCoef ci = load_axis(p.axI, i); cib_a = ci.cab; cib_b = ci.cbb; cif_a = ci.caf; cif_b = ci.cbf; }
            // labels of the pair and of its +j / +k / +i neighbours: three labels along k serve both cells
            const unsigned la = l0p[0], lb = l0p[1], lk2 = l0p[2];
            float b0[2], bi[2], bj[2], bk[2];
            {
                const unsigned m0[2] = { la & MSK, lb & MSK }, mi[2] = { l1p[0] & MSK, l1p[1] & MSK };
                const unsigned mj[2] = { l0p[LW] & MSK, l0p[LW + 1] & MSK }, mk[2] = { lb & MSK, lk2 & MSK };
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    if (SMC) { b0[c] = sB[m0[c]]; bi[c] = sB[mi[c]]; bj[c] = sB[mj[c]]; bk[c] = sB[mk[c]]; }
                    else { b0[c] = __ldg(&p.coef[m0[c]].B); bi[c] = __ldg(&p.coef[mi[c]].B); bj[c] = __ldg(&p.coef[mj[c]].B); bk[c] = __ldg(&p.coef[mk[c]].B); }
                }
            }
            float X[2][9];
            {
                const float2 y0 = ld2(byy), yp1 = ld2(byy + SW), yp2 = ld2(byy + 2 * SW), ym1 = ld2(byy - SW);
                const float2 z0 = ld2(bzz), z2 = ld2(bzz + 2);
                const float zk1 = bzz[-1];
                X[0][0] = D4C(cif_a, cif_b, xx_p1.x, xx_0.x, xx_p2.x, xx_m1.x);
                X[1][0] = D4C(cif_a, cif_b, xx_p1.y, xx_0.y, xx_p2.y, xx_m1.y);
                X[0][3] = D4C(cib_a, cib_b, xy_0.x, xy_m1.x, xy_p1.x, xy_m2.x);
                X[1][3] = D4C(cib_a, cib_b, xy_0.y, xy_m1.y, xy_p1.y, xy_m2.y);
                X[0][6] = D4C(cib_a, cib_b, xz_0.x, xz_m1.x, xz_p1.x, xz_m2.x);
                X[1][6] = D4C(cib_a, cib_b, xz_0.y, xz_m1.y, xz_p1.y, xz_m2.y);
                X[0][4] = D4C(cjf_a, cjf_b, yp1.x, y0.x, yp2.x, ym1.x);
                X[1][4] = D4C(cjf_a, cjf_b, yp1.y, y0.y, yp2.y, ym1.y);
                X[0][8] = D4C(ckf_a[0], ckf_b[0], z0.y, z0.x, z2.x, zk1);
                X[1][8] = D4C(ckf_a[1], ckf_b[1], z2.x, z0.y, z2.y, z0.x);
                if constexpr (SHEAR) {
                    const float2 a0 = ld2(bxy), am1 = ld2(bxy - SW), ap1 = ld2(bxy + SW), am2 = ld2(bxy - 2 * SW);
                    const float2 cm = ld2(bxz - 2), c0 = ld2(bxz);
                    const float c2 = bxz[2];
                    const float2 em = ld2(byz - 2), e0 = ld2(byz), em1 = ld2(byz - SW), ep1 = ld2(byz + SW), em2 = ld2(byz - 2 * SW);
                    const float e2 = byz[2];
                    X[0][1] = D4C(cjb_a, cjb_b, a0.x, am1.x, ap1.x, am2.x);
                    X[1][1] = D4C(cjb_a, cjb_b, a0.y, am1.y, ap1.y, am2.y);
                    X[0][2] = D4C(ckb_a[0], ckb_b[0], c0.x, cm.y, c0.y, cm.x);
                    X[1][2] = D4C(ckb_a[1], ckb_b[1], c0.y, c0.x, c2, cm.y);
                    X[0][5] = D4C(ckb_a[0], ckb_b[0], e0.x, em.y, e0.y, em.x);
                    X[1][5] = D4C(ckb_a[1], ckb_b[1], e0.y, e0.x, e2, em.y);
                    X[0][7] = D4C(cjb_a, cjb_b, e0.x, em1.x, ep1.x, em2.x);
                    X[1][7] = D4C(cjb_a, cjb_b, e0.y, em1.y, ep1.y, em2.y);
                } else {
#pragma unroll
                    for (int c = 0; c < 2; c++) { X[c][1] = X[c][2] = X[c][5] = X[c][7] = 0.f; }
                }
            }
            float v[2][3];
            {
                const float2 a = ld2(pb + (QB_V + 0) * BOXFR), b = ld2(pb + (QB_V + 1) * BOXFR), c = ld2(pb + (QB_V + 2) * BOXFR);
                v[0][0] = a.x; v[1][0] = a.y; v[0][1] = b.x; v[1][1] = b.y; v[0][2] = c.x; v[1][2] = c.y;
            }
            if (upd[0] && upd[1] && !cpml[0] && !cpml[1]) {
                // ---------------- both cells interior (the common case, uniform over the warp): straight-line code, so that
                // the two cells' dependency chains interleave
                float bx[2], by[2], bz[2];
#pragma unroll
                for (int c = 0; c < 2; c++) { bx[c] = 0.5f * (b0[c] + bi[c]); by[c] = 0.5f * (b0[c] + bj[c]); bz[c] = 0.5f * (b0[c] + bk[c]); }
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    if constexpr (SHEAR) {
                        v[c][0] += dt * bx[c] * (X[c][0] + X[c][1] + X[c][2]);
                        v[c][1] += dt * by[c] * (X[c][3] + X[c][4] + X[c][5]);
                        v[c][2] += dt * bz[c] * (X[c][6] + X[c][7] + X[c][8]);
                    } else {
                        v[c][0] += dt * bx[c] * X[c][0];
                        v[c][1] += dt * by[c] * (X[c][3] + X[c][4]);
                        v[c][2] += dt * bz[c] * (X[c][6] + X[c][8]);
                    }
                }
                const bool ra = (la & LabelTraits<LT>::REFL) != 0, rb = (lb & LabelTraits<LT>::REFL) != 0;
#pragma unroll
                for (int n = 0; n < 3; n++) { v[0][n] = ra ? 0.f : v[0][n]; v[1][n] = rb ? 0.f : v[1][n]; }
                if (ACC) {
#pragma unroll
                    for (int c = 0; c < 2; c++) {
                        const unsigned qa = q + c - 2 * s1;
                        accumulate(p, BB_MAP_VX, qa, v[c][0], false);
                        accumulate(p, BB_MAP_VY, qa, v[c][1], false);
                        accumulate(p, BB_MAP_VZ, qa, v[c][2], false);
                        accumulate(p, BB_MAP_ALLV, qa, v[c][0] * v[c][0] + v[c][1] * v[c][1] + v[c][2] * v[c][2], true);
                    }
                }
            } else {
#pragma unroll
            for (int c = 0; c < 2; c++) {
                if (!upd[c]) continue;
                const float bx = 0.5f * (b0[c] + bi[c]), by = 0.5f * (b0[c] + bj[c]), bz = 0.5f * (b0[c] + bk[c]);
                if (cpml[c]) {
                    PmlCell pcell;
                    pcell.xd = xd; pcell.jd = jd; pcell.kd = kd[c];
                    const int ipx = i < p.P ? i - p.i0 : p.nxlo + (i - p.xhi_begin);
                    pcell.qx = (unsigned)ipx * s1 + col + c; pcell.qy = qy + c; pcell.qz = qz[c];
                    pcell.cI = p.axI + i; pcell.cJ = sJ + ty; pcell.cK = sK + tx + c;
                    particle_pml<true>(p, pcell, bx, by, bz, X[c], v[c], pb + QB_X * BOXFR + c, pb + (yoff >> 2) + c, pzb + (po >> 2) + zsrc[c], BOXFR, zcomp);
                } else if constexpr (SHEAR) {
                    v[c][0] += dt * bx * (X[c][0] + X[c][1] + X[c][2]);
                    v[c][1] += dt * by * (X[c][3] + X[c][4] + X[c][5]);
                    v[c][2] += dt * bz * (X[c][6] + X[c][7] + X[c][8]);
                } else {
                    v[c][0] += dt * bx * X[c][0];
                    v[c][1] += dt * by * (X[c][3] + X[c][4]);
                    v[c][2] += dt * bz * (X[c][6] + X[c][8]);
                }
                if ((c == 0 ? la : lb) & LabelTraits<LT>::REFL) { v[c][0] = v[c][1] = v[c][2] = 0.f; }
                if (ACC && !cpml[c]) {
                    const unsigned qa = q + c - 2 * s1;
                    accumulate(p, BB_MAP_VX, qa, v[c][0], false);
                    accumulate(p, BB_MAP_VY, qa, v[c][1], false);
                    accumulate(p, BB_MAP_VZ, qa, v[c][2], false);
                    accumulate(p, BB_MAP_ALLV, qa, v[c][0] * v[c][0] + v[c][1] * v[c][1] + v[c][2] * v[c][2], true);
                }
            }
            }
            st2(p.V[0] + q, v[0][0], v[1][0]); st2(p.V[1] + q, v[0][1], v[1][1]); st2(p.V[2] + q, v[0][2], v[1][2]);
            if (pushsel) {
                if (p.bsrc_map) {      // particle sources of the boundary planes are injected here (the source kernel skips them)
                    const int bp = boundary_plane(p, pushsel, i);
                    bool any = false;
#pragma unroll
                    for (int c = 0; c < 2; c++) {
                        float val, ox, oy, oz;
                        if (bp >= 0 && active[c] && boundary_source(p, bp, col + c, val, ox, oy, oz)) {
                            if (p.src_hard) { v[c][0] = val * ox; v[c][1] = val * oy; v[c][2] = val * oz; }
                            else { v[c][0] += val * ox; v[c][1] += val * oy; v[c][2] += val * oz; }
                            any = true;
                        }
                    }
                    if (any) { st2(p.V[0] + q, v[0][0], v[1][0]); st2(p.V[1] + q, v[0][1], v[1][1]); st2(p.V[2] + q, v[0][2], v[1][2]); }
                }
                if ((pushsel & 1) && i < p.i0 + 2) {
                    float *b = p.peerV[0];
                    const unsigned qn = (p.peer_plane[0] + (unsigned)(i - p.i0)) * s1 + col;
                    st2(b + qn, v[0][0], v[1][0]); st2(b + p.peer_vol[0] + qn, v[0][1], v[1][1]); st2(b + 2 * p.peer_vol[0] + qn, v[0][2], v[1][2]);
                }
                if ((pushsel & 2) && i >= p.i1 - 2) {
                    float *b = p.peerV[1];
                    const unsigned qn = (p.peer_plane[1] + (unsigned)(i - (p.i1 - 2))) * s1 + col;
                    st2(b + qn, v[0][0], v[1][0]); st2(b + p.peer_vol[1] + qn, v[0][1], v[1][1]); st2(b + 2 * p.peer_vol[1] + qn, v[0][2], v[1][2]);
                }
            }
        };
        if (upd[0] || upd[1]) {
            if (fsh) pair_update(std::true_type{}); else pair_update(std::false_type{});
        }
        if (pushsel) {
            const bool last_lo = (pushsel & 1) && i == min(ic1, p.i0 + 2) - 1;
            const bool last_hi = (pushsel & 2) && i == ic1 - 1;
            if ((last_lo || last_hi) && p.publish) {
                asm volatile("bar.sync 1, %0;" ::"n"(G::NT) : "memory");
                if (tid == 0) {
                    __threadfence_system();
                    const unsigned expected = 2u * gridDim.x * gridDim.y;
                    if (last_lo) peer_publish(p, 0, (unsigned)(min(ic1, p.i0 + 2) - ic0), expected);
                    if (last_hi) peer_publish(p, 1, (unsigned)(ic1 - max(ic0, p.i1 - 2)), expected);
                }
            }
        }
        __syncwarp();
        if (lane0) { mbar_arrive(hb0 + MAX_NSH * 8); mbar_arrive(pbar + MAX_NSP * 8); }
        if constexpr (G::FOLD) {
            if (lane0 && warp == 0 && it + nsh < nplanes + 2) { load_halo(it + nsh, rh.slot, rh.par); rh.advance(); }
            if (lane0 && warp == 1 && it + nsp < nplanes) { load_point(it + nsp, rp.slot, rp.par); rp.advance(); }
        }
        ho = ho1; hb0 = hb1; ho1 = ho2; hb1 = hb2;
        ho2 += hstage; hb2 += 8;
        if (ho2 == hend) { ho2 = 0; hb2 = fullH; hpar ^= 1u; }
        po += pstage; pbar += 8;
        if (po == pend) { po = 0; pbar = fullP; ppar ^= 1u; }
    }
    if (p.dbg && tid == 0) {
        unsigned long long *d = p.dbg + 4ull * ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x);
        d[0] = dbg_t0; d[1] = globaltimer_ns(); d[2] = ((unsigned long long)blockIdx.z << 40) | ((unsigned long long)blockIdx.y << 20) | blockIdx.x; d[3] = (unsigned long long)np;
    }
}
}  // namespace tma
