// kernel_variant = 0: the production half-step kernels.
//
// One CTA owns an 8 x 64 (j,k) tile of the plane (k = the contiguous axis of the caller's C-order
// volumes) and marches `chunk` planes along the slab axis i.  Every operand of a plane reaches
// shared memory through TMA (cp.async.bulk.tensor, 3-D descriptors over the pitched volumes) into
// two mbarrier-tracked rings that run several planes ahead of the arithmetic:
//   * the "halo ring" holds the stencil inputs of the half-step (V for the stress kernel, the
//     stresses for the particle kernel) as (8+4) x (64+8) boxes plus the label box; out-of-volume
//     taps are zero-filled by the TMA unit, so the kernels carry no boundary branches for loads;
//   * the "point ring" holds the read-modify-write fields of the cell itself (stresses, memory
//     variables, pressure and the damped PML parts for the stress kernel; V and its damped parts
//     for the particle kernel) as 8 x 64 boxes.
// Warps 16 and 17 of the CTA are the producers of the halo ring and of the point ring (independent,
// so neither ring throttles the other's prefetch distance): one lane waits on the slot's "empty"
// mbarrier and issues the TMA loads of the next plane against the slot's "full" mbarrier.  Warps 0-15 are consumers (half a
// tile row each): they wait on "full", compute, store, and release the slot with one arrive per warp.
// There is no CTA-wide barrier inside the plane loop, so warps drift apart by up to the ring depth.
// The i-direction stencil lives in a register queue fed from the halo ring; the in-plane stencil
// reads the ring directly (row pitch 72 floats: conflict-free).  Results go straight from registers
// to global memory (one 128-byte row segment per warp and field, two warps side by side).
//
// Which fields move at all is decided per (plane, tile) from the flag byte computed once per
// simulation (flags_kernel): memory variables only where something attenuates, shear stresses only
// where something is solid, nothing but the split parts inside the PML.  The PML shell is handled
// in the same launch: a cell whose axis is damped updates the stored damped part of that axis and
// adds the increment to the total field (fdtd_cell.cuh).
#pragma once
#include <type_traits>
#include "fdtd_cell.cuh"

namespace tma {
// 8 x 64 tile: 256-byte row segments (profiles/tile_bw_probe.cu: 128-byte segments cap the access
// pattern at ~70% of the HBM copy bandwidth, 256-byte segments at ~90-98%)
constexpr int TX = 64, TY = BB_TY, HALO = 2;
// other tile heights (BB_TILE_ROWS 9, 10: +3 % when last measured) have not been re-validated since the compact ring stages
// and the k-PML thread remap went in -- builds with 10 and 12 rows faulted on a B200 -- so they do not compile for now
static_assert(TY == 8 || TY == 12, "tile heights other than 8 and 12 rows are not validated");
// the innermost TMA coordinate must be a multiple of 16 bytes (measured: a box starting at k0-2
// raises an illegal-instruction fault), so halo boxes start at k0-4 and are TX+8 floats wide
constexpr int HK = 4;
constexpr int SW = TX + 2 * HK;          // 72
constexpr int SH = TY + 2 * HALO;        // 12
constexpr int NT = TX * TY;              // 512 consumer threads, one cell each
constexpr int NCW = NT / 32;             // consumer warps
constexpr int CTAS_PER_SM = TY <= 4 ? 2 : 1;
constexpr int NTB = NT + 64;             // + two producer warps (halo ring, point ring)
constexpr int HBOX = SW * SH * 4;        // bytes per halo box; the boxes of one 4-D TMA land densely
constexpr int HBOX_STRIDE = HBOX;
__host__ __device__ constexpr int align128(int x) { return (x + 127) & ~127; }
constexpr int PBOX = TX * TY * 4;        // 2048 bytes per point box
constexpr int LH = TY + 1;               // label rows j0 .. j0+TY
constexpr int LBOX_STRIDE = ((TX + 8) * 2 * LH + 127) & ~127;   // >= LW*LH*sizeof(LT) for both label types, 128-byte aligned
// label box width: >= TX+1 labels and a multiple of 16 bytes: uint8 80 labels, uint16 72 labels
template <typename LT> struct LabBox { static constexpr int W = sizeof(LT) == 1 ? TX + 16 : TX + 8; };
constexpr int MAXCHUNK = 64;
constexpr int MAX_ZBW = 32;              // PML thickness supported by the compact Z-part boxes

// ---------------------------------------------------------------- PTX wrappers
// barriers and TMA destinations are addressed by 32-bit shared-window addresses computed once per thread
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// BB_WAIT_HINT_NS > 0: the try_wait may park the thread in hardware for up to that many nanoseconds before it reports
// "not yet", instead of the default (shorter) limit -- fewer trips round the spin loop for a warp that is ahead of its data
// BB_SKELETON (profiling builds only): 1 = waits and arrives without the cell update, 2 = + the loads / stores of a fluid cell
#ifndef BB_SKELETON
#define BB_SKELETON 0
#endif
#ifndef BB_WAIT_HINT_NS
#define BB_WAIT_HINT_NS 0
#endif
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
#if BB_WAIT_HINT_NS > 0
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"((uint32_t)BB_WAIT_HINT_NS) : "memory");
#else
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
#endif
    return ok != 0;
}
// a TMA that never lands (bad descriptor, wrong byte count) must not hang the GPU: trap after ~seconds
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) { if (++spins > (1u << 26)) __trap(); }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// ---- NVLink halo push (see DevParams::peerV): wait for the neighbours' previous half-step, publish this one
__device__ __forceinline__ void peer_wait(const DevParams &p, int side) {
    const unsigned long long want = p.seq - 1;                   // the neighbour's previous half-step of this epoch
    const volatile unsigned long long *f = p.flag_local + side;
    // a neighbour can be arbitrarily late (profiler replay, sanitizer, a shared GPU): wait on the wall clock, and on a
    // timeout flag the run as failed (bb_fdtd_run reports it) instead of trapping the context
    unsigned long long t0 = 0;
    unsigned spins = 0;
    while (*f < want) {
        __nanosleep(64);
        if ((++spins & 1023u) == 0 && p.peer_timeout_ns) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            if (!t0) t0 = t;
            else if (t - t0 > p.peer_timeout_ns) { atomicExch(p.err, 1 + side); break; }
        }
    }
    if (p.exp & 16) __threadfence_system(); else __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");             // the TMA unit reads what the neighbour wrote
}
// Called by one thread of a CTA (the halo-ring producer) once every consumer warp has released the side's last pushed
// plane: the mbarrier arrive / wait pair orders the warps' peer stores before this thread, its gpu-scope fence (issued by
// the caller) and the atomic count release them to the CTA that completes the count, and that CTA's system-scope fence --
// one per side and launch, cumulative over everything the counts released -- makes them visible to the neighbour GPU
// before the flag write.  (A system-scope fence in every boundary CTA cost 0.06-0.3 ms per half-step on the 1 MHz slabs:
// profiles/r2_scaling_experiments.txt.)
__device__ __forceinline__ void peer_publish(const DevParams &p, int side, unsigned planes_pushed, unsigned expected) {
    const unsigned before = atomicAdd(p.push_count + side, planes_pushed);
    if (before + planes_pushed == expected) {
        p.push_count[side] = 0;                                   // ready for the next launch
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(p.flag_peer[side]) = p.seq;
    }
}
// value and weights of the source (if any) sitting at cell `col` of boundary plane b of the slab (DevParams::bsrc_map)
__device__ __forceinline__ bool boundary_source(const DevParams &p, int b, unsigned col, float &value, float &ox, float &oy, float &oz) {
    const int sidx = p.bsrc_map[(long long)b * p.plane + col];
    if (sidx < 0) return false;
    const int row = p.bsrc_row[sidx];
    value = p.sf_row ? p.sf_row[row] : fmaf(p.env_sin, p.tone_ac[row], p.env_cos * p.tone_as[row]);
    ox = p.bsrc_o[0][sidx]; oy = p.bsrc_o[1][sidx]; oz = p.bsrc_o[2][sidx];
    return true;
}
// boundary plane index of plane i for the sides this CTA pushes (pushsel bit 0: lower, bit 1: upper), or -1
__device__ __forceinline__ int boundary_plane(const DevParams &p, int pushsel, int i) {
    if ((pushsel & 1) && i < p.i0 + 2) return i - p.i0;
    if ((pushsel & 2) && i >= p.i1 - 2) return 2 + i - (p.i1 - 2);
    return -1;
}
// keep a loop-invariant value in a register: the compiler otherwise re-derives it from the constant bank / special
// registers in every iteration of the plane loop (S2R + LDC + IMAD chains seen in the SASS)
// (ptxas does the re-deriving, so the value has to pass through an instruction it cannot see through: a shuffle from
// the thread's own lane, executed once per CTA)
__device__ __forceinline__ void keep(int &x) { x = __shfl_sync(0xffffffffu, x, threadIdx.x & 31); }
__device__ __forceinline__ void keep(unsigned &x) { x = __shfl_sync(0xffffffffu, x, threadIdx.x & 31); }
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// ---------------------------------------------------------------- tile flags
template <typename LT>
__global__ void __launch_bounds__(NT) flags_kernel(const DevParams p, unsigned char *__restrict__ flags) {
    __shared__ unsigned sflag;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
    const int k0 = blockIdx.x * TX, j0 = blockIdx.y * TY, ip = blockIdx.z;
    const int i = p.i0 - 2 + ip;
    if (tid == 0) sflag = 0;
    __syncthreads();
    const LT *lab = reinterpret_cast<const LT *>(p.lab);
    const unsigned MSK = LabelTraits<LT>::MASK;
    unsigned f = 0;
    const bool xd = in_pml1(i, p.n1, p.P);
    for (int t = tid; t < SH * SW; t += NT) {
        const int r = t / SW, c = t - r * SW;
        const int jj = j0 - HALO + r, kk = k0 - HK + c;
        if (jj < 0 || jj >= p.n2 || kk < 0 || kk >= p.n3 || c < HK - HALO || c >= HK + TX + HALO) continue;
        const long long q = ((long long)ip * p.n2 + jj) * p.pitch + kk;
        const unsigned m = lab[q] & MSK;
        const bool solid = __ldg(&p.coef[m].invG) < 3.0e38f;
        const bool inner = r >= HALO && r < HALO + TY && c >= HK && c < HK + TX;
        const bool ext1 = r >= HALO && r <= HALO + TY && c >= HK && c <= HK + TX;
        if (solid) f |= TF_SHEAR | (ext1 ? TF_SOLID : 0);
        if (ext1 && ip + 1 < p.nloc) {   // plane i+1 feeds the xy / xz edges of plane i
            const unsigned m1 = lab[q + p.plane] & MSK;
            if (__ldg(&p.coef[m1].invG) < 3.0e38f) f |= TF_SOLID;
        }
        if (inner && !xd && !in_pml1(jj, p.n2, p.P) && !in_pml1(kk, p.n3, p.P)) {
            f |= TF_INT;
            if (__ldg(&p.coef[m].tauS) != 0.0f || __ldg(&p.coef[m].LMCb) != 0.0f) f |= TF_ATT;
        }
    }
    if (f) atomicOr(&sflag, f);
    __syncthreads();
    if (tid == 0) {
        const unsigned plane_bits = (xd ? TF_XD : 0) | (i <= 1 || i >= p.n1 - 2 ? TF_IEDGE : 0) | (i >= p.n1 - 1 ? TF_ILAST : 0);
        flags[((long long)ip * p.ntj + blockIdx.y) * p.ntk + blockIdx.x] = (unsigned char)(sflag | plane_bits);
    }
}

// ---------------------------------------------------------------- ring bookkeeping
// position in a ring of `ns` slots: slot index and the parity the next wait on that slot must observe
struct RingPos {
    int slot, ns;
    unsigned par;
    __device__ RingPos(int nslots, int first_slot, unsigned parity) : slot(first_slot), ns(nslots), par(parity) {}
    __device__ void advance() { if (++slot == ns) { slot = 0; par ^= 1u; } }
};

// ---------------------------------------------------------------- shared-memory layout
// Every CTA gets the same dynamic allocation (one CTA per SM); what the stages of a CTA do not need (no Y parts away
// from the j-PML, no Z parts away from the k-PML, no shear boxes where nothing is solid, no X parts outside the i-PML)
// becomes deeper rings, see ring_depths().
constexpr int SMEM_BYTES = (CTAS_PER_SM == 1 ? 222 : 110) * 1024;
#ifndef BB_MIN_NSH
#define BB_MIN_NSH 6
#endif
constexpr int MAX_NSH = 12, MAX_NSP = 8, MIN_NSH = BB_MIN_NSH;
constexpr int OFF_COEF = 0;                                                      // MatCoef[128] (stress) / float B[128] (particle)
constexpr int OFF_AXJ = OFF_COEF + BB_MAX_SMEM_MAT * (int)sizeof(MatCoef);
constexpr int OFF_AXK = OFF_AXJ + TY * (int)sizeof(AxisCoef);
constexpr int OFF_FLAGS = OFF_AXK + TX * (int)sizeof(AxisCoef);
constexpr int OFF_BAR = OFF_FLAGS + ((MAXCHUNK + 8 + 15) / 16) * 16;
constexpr int OFF_RINGS = 9216;
static_assert(OFF_BAR + 2 * (MAX_NSH + MAX_NSP) * 8 <= OFF_RINGS, "tables overflow their 9 KB");
// ring depths of a CTA: as many point stages as fit beside MIN_NSH halo stages (the consumers hold three halo
// slots at a time, so that is three planes of halo prefetch), the rest of the shared memory as halo stages.
// (ncu, 4 point stages for every CTA: the consumers spent 8 % of their time waiting for the point ring.)
__device__ __forceinline__ void ring_depths(int pstage, int hstage, int &nsp, int &nsh) {
    constexpr int avail = SMEM_BYTES - OFF_RINGS;
    nsp = max(3, min(MAX_NSP, (avail - MIN_NSH * hstage) / pstage));
    nsh = min(MAX_NSH, (avail - nsp * pstage) / hstage);
    // the largest stages (solid tile inside two PML slabs, tall tiles): double-buffer the point ring rather than starve
    // the halo ring, of which the consumers hold three slots at a time
    if (nsh < 5) { nsp = 2; nsh = min(MAX_NSH, (avail - nsp * pstage) / hstage); }
}

// =========================================================================================
// stress half-step
// =========================================================================================
// point-box order inside a stage: the eight boxes every tile needs first, the shear stresses and their memory
// variables behind them -- a CTA whose planes hold nothing solid (TF_SOLID clear on all of them) leaves those six
// boxes out of its stages and gets deeper rings instead.  On a plane inside the i-PML the X parts use the five
// boxes from PB_RXX on (no interior cell exists on such a plane, so neither memory variables nor the pressure
// boxes are needed there); the Y / Z parts follow at full boxes, the Z parts as compact regions
enum { PB_SXX = 0, PB_SYY, PB_SZZ, PB_RXX, PB_RYY, PB_RZZ, PB_PR, PB_ACC, PB_FLUID, PB_SXY = PB_FLUID, PB_SXZ, PB_SYZ, PB_RXY, PB_RXZ, PB_RYZ, PB_PARTS };
constexpr int ST_LOFF = align128(3 * HBOX_STRIDE);         // label box behind the three V boxes
constexpr int ST_HSTAGE = ST_LOFF + LBOX_STRIDE;

template <typename LT, int ACC, bool PEER>
__global__ void __launch_bounds__(NTB, CTAS_PER_SM) stress_tma(const __grid_constant__ StressMaps tm, const DevParams p, const ChunkPlan plan) {
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
    // which damped parts this tile can need (CTA-uniform)
    const bool tile_jd = (int)blockIdx.y < p.nylo || (int)blockIdx.y >= p.tjhi0;
    const bool tile_zlo = k0 < p.P, tile_zhi = k0 + TX > p.n3 - p.P;
    const int yt = ((int)blockIdx.y < p.nylo ? (int)blockIdx.y : (int)blockIdx.y - p.tjhi0 + p.nylo) * TY;
    // point-stage layout (bytes): 14 full boxes, then the Y part boxes, then the compact Z part regions
    // (zbw columns x TY rows per component) of the low and the high side
    // per-plane traffic flags of this CTA's planes; any_solid = some plane of the chunk moves shear fields
    int fl = 0;
    if (tid < np + 2) {
        const int ipl = ipl0 + tid;
        fl = ipl < p.nloc ? p.flags[((long long)ipl * p.ntj + blockIdx.y) * p.ntk + blockIdx.x] : 0;
        sF[tid] = (unsigned char)fl;
    }
    const bool any_solid = __syncthreads_or(tid < np && (fl & TF_SOLID)) != 0;
    const int zcomp = p.zbw * TY;                                          // floats per Z part component
    const int zshear = align128(3 * zcomp * 4);                          // the two shear parts start 128-byte aligned (own TMA)
    const int zreg = zshear + (any_solid ? align128(2 * zcomp * 4) : 0);
    const int yoff = (any_solid ? PB_PARTS : PB_FLUID) * PBOX, zlo_off = yoff + (tile_jd ? (any_solid ? 5 : 3) * PBOX : 0);
    const int zhi_off = zlo_off + (tile_zlo ? zreg : 0);
    const int pstage = zhi_off + (tile_zhi ? zreg : 0);
    int nsp, nsh;
    ring_depths(pstage, ST_HSTAGE, nsp, nsh);
    const int offP = OFF_RINGS, offH = OFF_RINGS + nsp * pstage;

    // ---- per-CTA tables
    if (SMC) for (int t = tid; t < p.nmat * (int)(sizeof(MatCoef) / 4); t += NTB) reinterpret_cast<float *>(sC)[t] = reinterpret_cast<const float *>(p.coef)[t];
    if (tid < TY * 8) { const int r = tid >> 3, e = tid & 7; reinterpret_cast<float *>(sJ)[tid] = reinterpret_cast<const float *>(p.axJ + min(j0 + r, p.n2 - 1))[e]; }
    if (tid < TX * 8) { const int r = tid >> 3, e = tid & 7; reinterpret_cast<float *>(sK)[tid] = reinterpret_cast<const float *>(p.axK + min(k0 + r, p.n3 - 1))[e]; }
    // halo planes received through NVLink: the neighbour's previous half-step must have landed before this CTA
    // (its TMA loads and its queue prologue) reads them
    const bool first_hs = (unsigned)(p.seq & 0xffffffffu) <= 1u;
    const bool near_lo = PEER && ic0 < p.i0 + 2 && p.peerV[0] != nullptr, near_hi = PEER && ic1 > p.i1 - 2 && p.peerV[1] != nullptr;
    const bool has_peer = PEER && (ic0 < p.i0 + 2 || ic1 > p.i1 - 2) && (p.peerV[0] != nullptr || p.peerV[1] != nullptr);   // this CTA may push planes
    if (tid == 0) {
        if (!first_hs && near_lo && !(p.exp & 4)) peer_wait(p, 0);
        if (!first_hs && near_hi && !(p.exp & 4)) peer_wait(p, 1);
        for (int s = 0; s < nsh; s++) { mbar_init(fullH + s * 8, 1); mbar_init(emptyH + s * 8, NCW); }
        for (int s = 0; s < nsp; s++) { mbar_init(fullP + s * 8, 1); mbar_init(emptyP + s * 8, NCW); }
        fence_barrier_init();
    }
    __syncthreads();

    // =============================== producer warps ===============================
    if (warp == NCW) {          // halo ring: the three V boxes (one 4-D TMA) + labels of plane ic0 + r
        if (lane != 0) return;
        RingPos rh(nsh, 0, 1);
        // This thread also publishes the slab-boundary planes the CTA pushes to a neighbour: the wait for a halo slot to come
        // back tells it that every consumer warp has finished (stored, pushed, released) the plane that used the slot, so
        // the fence, the count and the flag write cost the consumer warps nothing (150-300 us per launch on the 1 MHz
        // slabs when consumer thread 0 did them behind a barrier: ~30 boundary CTAs per SM).
        const int lo_last = (ic0 < p.i0 + 2 && p.peerS[0]) ? min(ic1, p.i0 + 2) - 1 - ic0 : -1;
        const int hi_last = (ic1 > p.i1 - 2 && p.peerS[1]) ? np - 1 : -1;
        const int rend = (PEER && p.publish) ? max(np + 2, max(lo_last, hi_last) + nsh + 1) : np + 2;
        for (int r = 0; r < rend; r++) {
            const int slot = rh.slot;
            mbar_wait(emptyH + slot * 8, rh.par);
            if (PEER && p.publish && r >= nsh && (r - nsh == lo_last || r - nsh == hi_last)) {
                if (p.exp & 32) __threadfence_system(); else if (!(p.exp & 1)) __threadfence();
                const unsigned expected = 2u * gridDim.x * gridDim.y;
                if (r - nsh == lo_last) peer_publish(p, 0, (unsigned)(min(ic1, p.i0 + 2) - ic0), expected);
                if (r - nsh == hi_last) peer_publish(p, 1, (unsigned)(ic1 - max(ic0, p.i1 - 2)), expected);
            }
            rh.advance();
            if (r >= np + 2) continue;
            const uint32_t st = sm32 + offH + slot * ST_HSTAGE;
            const uint32_t bar = fullH + slot * 8;
            mbar_expect_tx(bar, 3 * HBOX + LW * LH * (int)sizeof(LT));
            const int ipl = ipl0 + r;
            tma_load_4d(st, &tm.v3, bar, k0 - HK, j0 - HALO, ipl, 0);
            tma_load_3d(st + ST_LOFF, &tm.lab, bar, k0, j0, ipl);
        }
        return;
    }
    if (warp == NCW + 1) {      // point ring: read-modify-write fields and damped parts of plane ic0 + r
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
    // thread -> cell (ty, tx) of the tile.  Plain tiles: a warp is half a row.  Tiles holding k-PML columns: the PML
    // cells of all rows are enumerated first, then the interior cells, so that a warp runs either the split-field
    // path or the interior path instead of both (8 rows x 12 PML columns are exactly three warps).
    int tx, ty;
    bool mapped = true;
    if (!(tile_zlo || tile_zhi)) { tx = tid & (TX - 1); ty = tid / TX; }
    else {
        const int wcols = min(TX, p.n3 - k0);                                   // columns of the tile inside the grid
        const int nlo = tile_zlo ? min(p.P - k0, wcols) : 0;                    // PML columns [0, nlo)
        const int hi0 = tile_zhi ? max(p.n3 - p.P - k0, nlo) : wcols;           // PML columns [hi0, wcols)
        const int npc = nlo + (wcols - hi0), nic = hi0 - nlo;
        if (tid < TY * npc) { ty = tid / npc; const int c = tid - ty * npc; tx = c < nlo ? c : hi0 + (c - nlo); }
        else if (tid - TY * npc < TY * nic) { const int t2 = tid - TY * npc; ty = t2 / nic; tx = nlo + (t2 - ty * nic); }
        else { tx = 0; ty = 0; mapped = false; }                                // columns beyond the grid: nothing to do
    }
    const int k = k0 + tx, j = j0 + ty;
    const bool active = mapped && k < p.n3 && j < p.n2;
    const bool jd = in_pml1(j, p.n2, p.P), kd = in_pml1(k, p.n3, p.P);
    const bool jkd = jd || kd;
    // which cells this thread updates: every in-grid cell outside the PML, split-field cells except those on the last
    // row / column / plane of the grid
    const bool upd_pml = active && j < p.n2 - 1 && k < p.n3 - 1;
    // difference coefficients of this thread's row and column (9/8, 1/24 away from the faces of the domain)
    const float cjb_a = sJ[ty].cab, cjb_b = sJ[ty].cbb, cjf_a = sJ[ty].caf, cjf_b = sJ[ty].cbf;
    const float ckb_a = sK[tx].cab, ckb_b = sK[tx].cbb, ckf_a = sK[tx].caf, ckf_b = sK[tx].cbf;
    unsigned s1 = (unsigned)p.plane;   // element indices fit 32 bits (checked at create)
    keep(s1);

    // ---- register queue along i (state before the shift of plane ic0)
    const float *__restrict__ Vx = p.V[0], *__restrict__ Vy = p.V[1], *__restrict__ Vz = p.V[2];
    const unsigned col = (unsigned)min(j, p.n2 - 1) * p.pitch + min(k, p.pitch - 1);
    unsigned q = (unsigned)ipl0 * s1 + col;
    float vx_m2, vx_m1 = Vx[q - 2 * s1], vx_0 = Vx[q - s1], vx_p1 = 0.f;
    float vy_m1, vy_0 = Vy[q - s1], vy_p1 = 0.f, vy_p2 = 0.f;
    float vz_m1, vz_0 = Vz[q - s1], vz_p1 = 0.f, vz_p2 = 0.f;
    const int sc = (ty + HALO) * SW + tx + HK;     // this thread's cell in a halo box
    const int lc = ty * LW + tx;                     // ... in a label box
    const int pc = ty * TX + tx;                     // ... in a point box
    const float dt = p.dt;
    const unsigned MSK = LabelTraits<LT>::MASK;
    // part arrays: index of this cell on plane ic0 and the per-plane strides
    unsigned qy_stride = (unsigned)p.nyrows * p.pitch, qz_stride = (unsigned)p.n2 * p.zpw;
    // (not pinned: only split-field cells use them, and pinning them made ptxas spill a loop counter to local memory)
    // boundary planes this CTA pushes to the slab neighbours: bit 0 lower, bit 1 upper (0 for almost every CTA)
    // (PEER = false, a slab without neighbours: a constant, and every block it guards leaves the instruction stream)
    int pushsel = 0;
    if constexpr (PEER) { pushsel = has_peer ? ((ic0 < p.i0 + 2 && p.peerS[0] ? 1 : 0) | (ic1 > p.i1 - 2 && p.peerS[1] ? 2 : 0)) : 0; keep(pushsel); }
    int nplanes = np, lane0 = lane == 0;
    keep(nplanes); keep(lane0);
    unsigned qy = ((unsigned)(ic0 - p.i0) * p.nyrows + yt + ty) * p.pitch + k;
    const int kz = k < p.P ? k : k - (p.n3 - p.P);                       // column inside the Z part box of this cell's side
    unsigned qz = ((unsigned)(ic0 - p.i0) * p.n2 + min(j, p.n2 - 1)) * p.zpw + (k < p.P ? kz : p.zbw + kz);
    const int zsrc = ((k < p.P ? zlo_off : zhi_off) >> 2) + ty * p.zbw + (kd ? kz : 0);   // float offset inside a point stage

    // per-thread pointers to this thread's own cell in slot 0 of each ring (kept in registers: everything inside
    // the loop is one of these plus a running byte offset)
    const char *hsc = reinterpret_cast<const char *>(sm + offH) + sc * 4;                        // halo boxes
    const char *lsc = reinterpret_cast<const char *>(sm + offH + ST_LOFF) + lc * sizeof(LT);   // label box
    const char *psc = reinterpret_cast<const char *>(sm + offP) + pc * 4;                        // point boxes
    const char *pzs = reinterpret_cast<const char *>(sm + offP) + zsrc * 4;                      // Z part region
    auto hbox = [&](int off, int c) { return reinterpret_cast<const float *>(hsc + off + c * HBOX_STRIDE); };
    auto lbox = [&](int off) { return reinterpret_cast<const LT *>(lsc + off); };
    // ring state as byte offsets / barrier addresses (no multiplies in the loop): halo slots of planes it, it+1,
    // it+2 (the one waited on inside the loop) and the point slot of plane it
    int ho = 0, ho1 = ST_HSTAGE, ho2 = 2 * ST_HSTAGE, po = 0;
    uint32_t hb0 = fullH, hb1 = fullH + 8, hb2 = fullH + 16, pbar = fullP;
    unsigned hpar = 0, ppar = 0;
    const int hend = nsh * ST_HSTAGE, pend = nsp * pstage;

    // planes ic0 and ic0+1 feed the queue before the loop
    mbar_wait(fullH, 0);
    vx_p1 = hbox(0, 0)[0]; vy_p1 = hbox(0, 1)[0]; vz_p1 = hbox(0, 2)[0];
    mbar_wait(fullH + 8, 0);
    vy_p2 = hbox(ST_HSTAGE, 1)[0]; vz_p2 = hbox(ST_HSTAGE, 2)[0];

    for (int it = 0; it < nplanes; it++, q += s1, qy += qy_stride, qz += qz_stride) {
        const int i = ic0 + it;
        const unsigned f = sF[it];
        mbar_wait(hb2, hpar);
        // ---------------- shift the queue: plane i becomes the centre
        vx_m2 = vx_m1; vx_m1 = vx_0; vx_0 = vx_p1; vx_p1 = hbox(ho1, 0)[0];
        vy_m1 = vy_0; vy_0 = vy_p1; vy_p1 = vy_p2; vy_p2 = hbox(ho2, 1)[0];
        vz_m1 = vz_0; vz_0 = vz_p1; vz_p1 = vz_p2; vz_p2 = hbox(ho2, 2)[0];
        mbar_wait(pbar, ppar);
        const bool xd = (f & TF_XD) != 0;
        const bool cellpml = xd || jkd;
        // the cell update, compiled twice: for planes of the tile where shear fields move (TF_SOLID) and for the others,
        // where every shear term (differences, rigidities, zero-filled registers, stores) is absent from the instruction stream
        auto cell_update = [&](auto solid_tag) {
            constexpr bool SOL = decltype(solid_tag)::value;
            const float *bx = hbox(ho, 0), *by = hbox(ho, 1), *bz = hbox(ho, 2);
            const LT *l0p = lbox(ho), *l1p = lbox(ho1);
            const float *pb = reinterpret_cast<const float *>(psc + po);
            const unsigned l0 = l0p[0];
            const bool refl = (l0 & LabelTraits<LT>::REFL) != 0;
            MatCoef c;
            if (SMC) c = sC[l0 & MSK]; else c = load_coef_global(p.coef, l0 & MSK);
            // ---------------- the nine staggered differences
            // staggered differences with the domain-edge rules folded into coefficients: j/k per thread (hoisted out of
            // the loop), i per plane (uniform) -- one code path for every cell
            float cib_a = BB_CA, cib_b = BB_CB, cif_a = BB_CA, cif_b = BB_CB;
            if (f & TF_IEDGE) { const AxisCoef ci = load_axis(p.axI, i); cib_a = ci.cab; cib_b = ci.cbb; cif_a = ci.caf; cif_b = ci.cbf; }
            float D[9];
            D[0] = D4C(cib_a, cib_b, vx_0, vx_m1, vx_p1, vx_m2);
            D[1] = D4C(cjb_a, cjb_b, by[0], by[-SW], by[SW], by[-2 * SW]);
            D[2] = D4C(ckb_a, ckb_b, bz[0], bz[-1], bz[1], bz[-2]);
            if constexpr (SOL) {
                D[3] = D4C(cif_a, cif_b, vy_p1, vy_0, vy_p2, vy_m1);
                D[4] = D4C(cjf_a, cjf_b, bx[SW], bx[0], bx[2 * SW], bx[-SW]);
                D[5] = D4C(cif_a, cif_b, vz_p1, vz_0, vz_p2, vz_m1);
                D[6] = D4C(ckf_a, ckf_b, bx[1], bx[0], bx[2], bx[-1]);
                D[7] = D4C(cjf_a, cjf_b, bz[SW], bz[0], bz[2 * SW], bz[-SW]);
                D[8] = D4C(ckf_a, ckf_b, by[1], by[0], by[2], by[-1]);
            } else { D[3] = D[4] = D[5] = D[6] = D[7] = D[8] = 0.f; }
            // ---------------- edge rigidities (only where something is solid)
            float rigxy = 0.f, rigxz = 0.f, rigyz = 0.f, texy = 0.f, texz = 0.f, teyz = 0.f;
            if constexpr (SOL) {
                const unsigned mi = l1p[0] & MSK, mj = l0p[LW] & MSK, mk = l0p[1] & MSK;
                const unsigned mij = l1p[LW] & MSK, mik = l1p[1] & MSK, mjk = l0p[LW + 1] & MSK;
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
                rigxy = rigidity4(c.invG, igi, igj, igij);
                rigxz = rigidity4(c.invG, igi, igk, igik);
                rigyz = rigidity4(c.invG, igj, igk, igjk);
                texy = 0.25f * (c.tauS + ti + tj + tij);
                texz = 0.25f * (c.tauS + ti + tk + tik);
                teyz = 0.25f * (c.tauS + tj + tk + tjk);
            }
            float s[6];
            s[0] = pb[PB_SXX * NT]; s[1] = pb[PB_SYY * NT]; s[2] = pb[PB_SZZ * NT];
            if constexpr (SOL) { s[3] = pb[PB_SXY * NT]; s[4] = pb[PB_SXZ * NT]; s[5] = pb[PB_SYZ * NT]; }
            else { s[3] = s[4] = s[5] = 0.f; }
            if (cellpml) {
                // ---------------- PML shell: damped split parts (old values staged by TMA)
                PmlCell pcell;
                pcell.xd = xd; pcell.jd = jd; pcell.kd = kd;
                const int ipx = i < p.P ? i - p.i0 : p.nxlo + (i - p.xhi_begin);
                pcell.qx = (unsigned)ipx * s1 + col; pcell.qy = qy; pcell.qz = qz;
                pcell.cI = p.axI + i; pcell.cJ = sJ + ty; pcell.cK = sK + tx;
                stress_pml<true>(p, pcell, c.M, c.L, rigxy, rigxz, rigyz, D, s, pb + PB_RXX * NT, pb + (yoff >> 2), reinterpret_cast<const float *>(pzs + po),
                                 reinterpret_cast<const float *>(pzs + po + zshear), NT, zcomp);
                if (refl) { s[0] = s[1] = s[2] = s[3] = s[4] = s[5] = 0.f; }
                p.S[0][q] = s[0]; p.S[1][q] = s[1]; p.S[2][q] = s[2];
                if constexpr (SOL) { p.S[3][q] = s[3]; p.S[4][q] = s[4]; p.S[5][q] = s[5]; }
            } else {
                // ---------------- interior: viscoelastic update
                const bool att = attenuates(c);
                float pr = pb[PB_PR * NT];
                float r0 = 0.f, r1 = 0.f, r2 = 0.f;
                if (att) { r0 = pb[PB_RXX * NT]; r1 = pb[PB_RYY * NT]; r2 = pb[PB_RZZ * NT]; }
                stress_normal_interior(c, dt, att, D[0], D[1], D[2], s[0], s[1], s[2], r0, r1, r2, pr);
                if (refl) { s[0] = s[1] = s[2] = 0.f; pr = 0.f; }
                p.S[0][q] = s[0]; p.S[1][q] = s[1]; p.S[2][q] = s[2]; p.Pr[q] = pr;
                if (att) { p.R[0][q] = r0; p.R[1][q] = r1; p.R[2][q] = r2; }
                if constexpr (SOL) {
                    if (rigxy != 0.f) {
                        float r = pb[PB_RXY * NT];
                        stress_shear_interior(c, dt, rigxy, texy, D[3] + D[4], s[3], r);
                        if (refl) s[3] = 0.f;
                        p.S[3][q] = s[3];
                        if (texy != 0.f) p.R[3][q] = r;
                    }
                    if (rigxz != 0.f) {
                        float r = pb[PB_RXZ * NT];
                        stress_shear_interior(c, dt, rigxz, texz, D[5] + D[6], s[4], r);
                        if (refl) s[4] = 0.f;
                        p.S[4][q] = s[4];
                        if (texz != 0.f) p.R[4][q] = r;
                    }
                    if (rigyz != 0.f) {
                        float r = pb[PB_RYZ * NT];
                        stress_shear_interior(c, dt, rigyz, teyz, D[7] + D[8], s[5], r);
                        if (refl) s[5] = 0.f;
                        p.S[5][q] = s[5];
                        if (teyz != 0.f) p.R[5][q] = r;
                    }
                }
                if (ACC == 1) {
                    const float v = -c.K * pr;
                    p.acc_rms[q - 2 * s1] = pb[PB_ACC * NT] + v * v;
                } else if (ACC == 2) {
                    const unsigned qa = q - 2 * s1;
#pragma unroll
                    for (int n = 0; n < 6; n++) accumulate(p, BB_MAP_SXX + n, qa, s[n], false);
                    accumulate(p, BB_MAP_PRESSURE, qa, -c.K * pr, false);
                }
            }
            // ---------------- boundary planes also go to the slab neighbour (the stresses its particle update differentiates along i)
            if (pushsel) {
            if (p.bsrc_map) {      // stress sources of the boundary planes are injected here (the source kernel skips them)
                const int bp = boundary_plane(p, pushsel, i);
                float val, ox, oy, oz;
                if (bp >= 0 && boundary_source(p, bp, col, val, ox, oy, oz)) {
                    const float w = val * ox;
                    if (p.src_hard) { s[0] = w; s[1] = w; s[2] = w; } else { s[0] += w; s[1] += w; s[2] += w; }
                    p.S[0][q] = s[0]; p.S[1][q] = s[1]; p.S[2][q] = s[2];
                }
            }
            if ((pushsel & 1) && i < p.i0 + 2) {
                float *b = p.peerS[0];
                const unsigned qn = (p.peer_plane[0] + (unsigned)(i - p.i0)) * s1 + col;
                b[qn] = s[0];
                if constexpr (SOL) { b[3 * p.peer_vol[0] + qn] = s[3]; b[4 * p.peer_vol[0] + qn] = s[4]; }
            }
            if ((pushsel & 2) && i >= p.i1 - 2) {
                float *b = p.peerS[1];
                const unsigned qn = (p.peer_plane[1] + (unsigned)(i - (p.i1 - 2))) * s1 + col;
                b[qn] = s[0];
                if constexpr (SOL) { b[3 * p.peer_vol[1] + qn] = s[3]; b[4 * p.peer_vol[1] + qn] = s[4]; }
            }
            }
        };
#if BB_SKELETON == 0
        if (cellpml ? (upd_pml && !(f & TF_ILAST)) : active) {
            if (f & TF_SOLID) cell_update(std::true_type{}); else cell_update(std::false_type{});
        }
#elif BB_SKELETON == 2      // profiling aid: the data movement of an attenuating-fluid cell without its arithmetic
        if (active) {
            const float *pb = reinterpret_cast<const float *>(psc + po);
            const float e = hbox(ho, 1)[0] + hbox(ho, 2)[1] + vx_0;
            p.S[0][q] = pb[PB_SXX * NT] + e; p.S[1][q] = pb[PB_SYY * NT]; p.S[2][q] = pb[PB_SZZ * NT];
            if (!cellpml) { p.R[0][q] = pb[PB_RXX * NT]; p.R[1][q] = pb[PB_RYY * NT]; p.R[2][q] = pb[PB_RZZ * NT]; p.Pr[q] = pb[PB_PR * NT]; }
        }
#else
        (void)cell_update;
#endif
        // ---------------- this warp is done with the slots of plane i
        __syncwarp();
        if (lane0) { mbar_arrive(hb0 + MAX_NSH * 8); mbar_arrive(pbar + MAX_NSP * 8); }
        ho = ho1; hb0 = hb1; ho1 = ho2; hb1 = hb2;
        ho2 += ST_HSTAGE; hb2 += 8;
        if (ho2 == hend) { ho2 = 0; hb2 = fullH; hpar ^= 1u; }
        po += pstage; pbar += 8;
        if (po == pend) { po = 0; pbar = fullP; ppar ^= 1u; }
    }
    if (p.dbg && tid == 0) {    // tid 0 is a consumer thread: it leaves the loop when the CTA's last plane is done
        unsigned long long *d = p.dbg + 4ull * ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x);
        d[0] = dbg_t0; d[1] = globaltimer_ns(); d[2] = ((unsigned long long)blockIdx.z << 40) | ((unsigned long long)blockIdx.y << 20) | blockIdx.x; d[3] = (unsigned long long)np;
    }
}

// =========================================================================================
// particle half-step
// =========================================================================================
// halo-stage boxes: Syy Szz (halo boxes), Sxx (point box, i-stencil only), labels, then the shear stresses Sxy Sxz Syz
// (halo boxes) -- a CTA whose planes carry no shear stress (TF_SHEAR clear on all of them) ends its stages before them;
// point-stage boxes: V (3), X parts (3, only when the chunk reaches into the i-PML), then the Y / Z parts the tile
// needs (3 each)
enum { HB_SYY = 0, HB_SZZ, HB_SXY, HB_SXZ, HB_SYZ };
enum { QB_V = 0, QB_X = 3, QB_PARTS = 6 };
constexpr int PT_XOFF = align128(2 * HBOX_STRIDE);         // Sxx point box behind Syy, Szz
constexpr int PT_LOFF = PT_XOFF + PBOX;
constexpr int PT_S3OFF = align128(PT_LOFF + LBOX_STRIDE);  // the three shear boxes (own TMA)
constexpr int PT_HSTAGE = PT_S3OFF + align128(3 * HBOX_STRIDE);

template <typename LT, int ACC, bool PEER>
__global__ void __launch_bounds__(NTB, CTAS_PER_SM) particle_tma(const __grid_constant__ ParticleMaps tm, const DevParams p, const ChunkPlan plan) {
    constexpr bool SMC = sizeof(LT) == 1;
    constexpr int LW = LabBox<LT>::W;
    extern __shared__ __align__(1024) unsigned char sm[];
    float *sB = reinterpret_cast<float *>(sm + OFF_COEF);
    AxisCoef *sJ = reinterpret_cast<AxisCoef *>(sm + OFF_AXJ);
    AxisCoef *sK = reinterpret_cast<AxisCoef *>(sm + OFF_AXK);
    unsigned char *sF = sm + OFF_FLAGS;
    const uint32_t sm32 = smem_u32(sm);
    const uint32_t fullH = sm32 + OFF_BAR, emptyH = fullH + MAX_NSH * 8, fullP = emptyH + MAX_NSH * 8, emptyP = fullP + MAX_NSP * 8;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k0 = blockIdx.x * TX, j0 = blockIdx.y * TY;
    const int ic0 = plan.start[blockIdx.z], ic1 = plan.end[blockIdx.z];
    const int np = ic1 - ic0;
    const unsigned long long dbg_t0 = (p.dbg && tid == 0) ? globaltimer_ns() : 0ull;
    const int ipl0 = ic0 - p.i0 + 2;
    const bool tile_jd = (int)blockIdx.y < p.nylo || (int)blockIdx.y >= p.tjhi0;
    const bool tile_zlo = k0 < p.P, tile_zhi = k0 + TX > p.n3 - p.P;
    const int yt = ((int)blockIdx.y < p.nylo ? (int)blockIdx.y : (int)blockIdx.y - p.tjhi0 + p.nylo) * TY;
    // point-stage layout (bytes): 6 full boxes, then the Y part boxes, then the compact Z part regions
    // (zbw columns x TY rows per component) of the low and the high side
    // per-plane traffic flags of this CTA's planes; any_shear = some plane the chunk reads carries shear stresses
    int fl = 0;
    if (tid < np + 2) {
        const int ipl = ipl0 + tid;
        fl = ipl < p.nloc ? p.flags[((long long)ipl * p.ntj + blockIdx.y) * p.ntk + blockIdx.x] : 0;
        sF[tid] = (unsigned char)fl;
    }
    const bool any_shear = __syncthreads_or(fl & TF_SHEAR) != 0;
    const bool any_xd = ic0 < p.P || ic1 > p.n1 - p.P;                    // the chunk holds planes of the i-PML
    const int hstage = any_shear ? PT_HSTAGE : PT_S3OFF;
    const int zcomp = p.zbw * TY;                                          // floats per Z part component
    const int zreg = align128(3 * zcomp * 4);
    const int yoff = (any_xd ? QB_PARTS : QB_X) * PBOX, zlo_off = yoff + (tile_jd ? 3 * PBOX : 0), zhi_off = zlo_off + (tile_zlo ? zreg : 0);
    const int pstage = zhi_off + (tile_zhi ? zreg : 0);
    int nsp, nsh;
    ring_depths(pstage, hstage, nsp, nsh);
    const int offP = OFF_RINGS, offH = OFF_RINGS + nsp * pstage;

    if (SMC) for (int t = tid; t < p.nmat; t += NTB) sB[t] = p.coef[t].B;
    if (tid < TY * 8) { const int r = tid >> 3, e = tid & 7; reinterpret_cast<float *>(sJ)[tid] = reinterpret_cast<const float *>(p.axJ + min(j0 + r, p.n2 - 1))[e]; }
    if (tid < TX * 8) { const int r = tid >> 3, e = tid & 7; reinterpret_cast<float *>(sK)[tid] = reinterpret_cast<const float *>(p.axK + min(k0 + r, p.n3 - 1))[e]; }
    // halo planes received through NVLink: the neighbour's previous half-step must have landed before this CTA
    // (its TMA loads and its queue prologue) reads them
    const bool first_hs = (unsigned)(p.seq & 0xffffffffu) <= 1u;
    const bool near_lo = PEER && ic0 < p.i0 + 2 && p.peerS[0] != nullptr, near_hi = PEER && ic1 > p.i1 - 2 && p.peerS[1] != nullptr;
    const bool has_peer = PEER && (ic0 < p.i0 + 2 || ic1 > p.i1 - 2) && (p.peerV[0] != nullptr || p.peerV[1] != nullptr);   // this CTA may push planes
    if (tid == 0) {
        if (!first_hs && near_lo && !(p.exp & 4)) peer_wait(p, 0);
        if (!first_hs && near_hi && !(p.exp & 4)) peer_wait(p, 1);
        for (int s = 0; s < nsh; s++) { mbar_init(fullH + s * 8, 1); mbar_init(emptyH + s * 8, NCW); }
        for (int s = 0; s < nsp; s++) { mbar_init(fullP + s * 8, 1); mbar_init(emptyP + s * 8, NCW); }
        fence_barrier_init();
    }
    __syncthreads();

    // =============================== producer warps ===============================
    if (warp == NCW) {          // halo ring: stresses with halo + Sxx + labels of plane ic0 + r
        if (lane != 0) return;
        RingPos rh(nsh, 0, 1);
        // publishes the pushed boundary planes once every consumer warp has released them (see stress_tma)
        const int lo_last = (ic0 < p.i0 + 2 && p.peerV[0]) ? min(ic1, p.i0 + 2) - 1 - ic0 : -1;
        const int hi_last = (ic1 > p.i1 - 2 && p.peerV[1]) ? np - 1 : -1;
        const int rend = (PEER && p.publish) ? max(np + 2, max(lo_last, hi_last) + nsh + 1) : np + 2;
        for (int r = 0; r < rend; r++) {
            const int slot = rh.slot;
            mbar_wait(emptyH + slot * 8, rh.par);
            if (PEER && p.publish && r >= nsh && (r - nsh == lo_last || r - nsh == hi_last)) {
                if (p.exp & 32) __threadfence_system(); else if (!(p.exp & 1)) __threadfence();
                const unsigned expected = 2u * gridDim.x * gridDim.y;
                if (r - nsh == lo_last) peer_publish(p, 0, (unsigned)(min(ic1, p.i0 + 2) - ic0), expected);
                if (r - nsh == hi_last) peer_publish(p, 1, (unsigned)(ic1 - max(ic0, p.i1 - 2)), expected);
            }
            rh.advance();
            if (r >= np + 2) continue;
            const uint32_t st = sm32 + offH + slot * hstage;
            const uint32_t bar = fullH + slot * 8;
            const bool fsh = sF[r] & TF_SHEAR;
            mbar_expect_tx(bar, (fsh ? 5 : 2) * HBOX + PBOX + LW * LH * (int)sizeof(LT));
            const int ipl = ipl0 + r;
            tma_load_4d(st, &tm.sh2, bar, k0 - HK, j0 - HALO, ipl, 1);
            if (fsh) tma_load_4d(st + PT_S3OFF, &tm.sh3, bar, k0 - HK, j0 - HALO, ipl, 3);
            tma_load_4d(st + PT_XOFF, &tm.sxx, bar, k0, j0, ipl, 0);
            tma_load_3d(st + PT_LOFF, &tm.lab, bar, k0, j0, ipl);
        }
        return;
    }
    if (warp == NCW + 1) {      // point ring: V and its damped parts of plane ic0 + r
        if (lane != 0) return;
        RingPos rp(nsp, 0, 1);
        for (int r = 0; r < np; r++) {
            const int slot = rp.slot;
            mbar_wait(emptyP + slot * 8, rp.par);
            const uint32_t st = sm32 + offP + slot * pstage;
            const uint32_t bar = fullP + slot * 8;
            const int i = ic0 + r, ipl = ipl0 + r, io = i - p.i0;
            const bool xd = in_pml1(i, p.n1, p.P);
            mbar_expect_tx(bar, (3 + (xd ? 3 : 0) + (tile_jd ? 3 : 0)) * PBOX + ((tile_zlo ? 3 : 0) + (tile_zhi ? 3 : 0)) * zcomp * 4);
            tma_load_4d(st + QB_V * PBOX, &tm.v3, bar, k0, j0, ipl, 0);
            if (xd) tma_load_4d(st + QB_X * PBOX, &tm.xp3, bar, k0, j0, i < p.P ? io : p.nxlo + (i - p.xhi_begin), 5);
            if (tile_jd) tma_load_4d(st + yoff, &tm.yp3, bar, k0, yt, io, 5);
            if (tile_zlo) tma_load_4d(st + zlo_off, &tm.zp3, bar, 0, j0, io, 5);
            if (tile_zhi) tma_load_4d(st + zhi_off, &tm.zp3, bar, p.zbw, j0, io, 5);
            rp.advance();
        }
        return;
    }

    // =============================== consumer warps ===============================
    // thread -> cell (ty, tx) of the tile.  Plain tiles: a warp is half a row.  Tiles holding k-PML columns: the PML
    // cells of all rows are enumerated first, then the interior cells, so that a warp runs either the split-field
    // path or the interior path instead of both (8 rows x 12 PML columns are exactly three warps).
    int tx, ty;
    bool mapped = true;
    if (!(tile_zlo || tile_zhi)) { tx = tid & (TX - 1); ty = tid / TX; }
    else {
        const int wcols = min(TX, p.n3 - k0);                                   // columns of the tile inside the grid
        const int nlo = tile_zlo ? min(p.P - k0, wcols) : 0;                    // PML columns [0, nlo)
        const int hi0 = tile_zhi ? max(p.n3 - p.P - k0, nlo) : wcols;           // PML columns [hi0, wcols)
        const int npc = nlo + (wcols - hi0), nic = hi0 - nlo;
        if (tid < TY * npc) { ty = tid / npc; const int c = tid - ty * npc; tx = c < nlo ? c : hi0 + (c - nlo); }
        else if (tid - TY * npc < TY * nic) { const int t2 = tid - TY * npc; ty = t2 / nic; tx = nlo + (t2 - ty * nic); }
        else { tx = 0; ty = 0; mapped = false; }                                // columns beyond the grid: nothing to do
    }
    const int k = k0 + tx, j = j0 + ty;
    const bool active = mapped && k < p.n3 && j < p.n2;
    const bool jd = in_pml1(j, p.n2, p.P), kd = in_pml1(k, p.n3, p.P);
    const bool jkd = jd || kd;
    // which cells this thread updates: every in-grid cell outside the PML, split-field cells except those on the last
    // row / column / plane of the grid
    const bool upd_pml = active && j < p.n2 - 1 && k < p.n3 - 1;
    // difference coefficients of this thread's row and column (9/8, 1/24 away from the faces of the domain)
    const float cjb_a = sJ[ty].cab, cjb_b = sJ[ty].cbb, cjf_a = sJ[ty].caf, cjf_b = sJ[ty].cbf;
    const float ckb_a = sK[tx].cab, ckb_b = sK[tx].cbb, ckf_a = sK[tx].caf, ckf_b = sK[tx].cbf;
    unsigned s1 = (unsigned)p.plane;   // element indices fit 32 bits (checked at create)
    keep(s1);

    // queues: Sxx holds i-1..i+2 ; Sxy, Sxz hold i-2..i+1 (state before the shift of plane ic0)
    const float *__restrict__ Sxx = p.S[0], *__restrict__ Sxy = p.S[3], *__restrict__ Sxz = p.S[4];
    const unsigned col = (unsigned)min(j, p.n2 - 1) * p.pitch + min(k, p.pitch - 1);
    unsigned q = (unsigned)ipl0 * s1 + col;
    float xx_m1, xx_0 = Sxx[q - s1], xx_p1 = 0.f, xx_p2 = 0.f;
    float xy_m2, xy_m1 = Sxy[q - 2 * s1], xy_0 = Sxy[q - s1], xy_p1 = 0.f;
    float xz_m2, xz_m1 = Sxz[q - 2 * s1], xz_0 = Sxz[q - s1], xz_p1 = 0.f;
    const int sc = (ty + HALO) * SW + tx + HK;
    const int lc = ty * LW + tx;
    const int pc = ty * TX + tx;
    const float dt = p.dt;
    const unsigned MSK = LabelTraits<LT>::MASK;
    unsigned qy_stride = (unsigned)p.nyrows * p.pitch, qz_stride = (unsigned)p.n2 * p.zpw;
    keep(qy_stride); keep(qz_stride);
    // boundary planes this CTA pushes to the slab neighbours: bit 0 lower, bit 1 upper (0 for almost every CTA)
    // (PEER = false, a slab without neighbours: a constant, and every block it guards leaves the instruction stream)
    int pushsel = 0;
    if constexpr (PEER) { pushsel = has_peer ? ((ic0 < p.i0 + 2 && p.peerV[0] ? 1 : 0) | (ic1 > p.i1 - 2 && p.peerV[1] ? 2 : 0)) : 0; keep(pushsel); }
    int nplanes = np, lane0 = lane == 0;
    keep(nplanes); keep(lane0);
    unsigned qy = ((unsigned)(ic0 - p.i0) * p.nyrows + yt + ty) * p.pitch + k;
    const int kz = k < p.P ? k : k - (p.n3 - p.P);                       // column inside the Z part box of this cell's side
    unsigned qz = ((unsigned)(ic0 - p.i0) * p.n2 + min(j, p.n2 - 1)) * p.zpw + (k < p.P ? kz : p.zbw + kz);
    const int zsrc = ((k < p.P ? zlo_off : zhi_off) >> 2) + ty * p.zbw + (kd ? kz : 0);   // float offset inside a point stage

    const char *hsc = reinterpret_cast<const char *>(sm + offH) + sc * 4;                                    // halo boxes
    const char *xsc = reinterpret_cast<const char *>(sm + offH + PT_XOFF) + pc * 4;                  // Sxx point box
    const char *lsc = reinterpret_cast<const char *>(sm + offH + PT_LOFF) + lc * sizeof(LT);  // label box
    const char *psc = reinterpret_cast<const char *>(sm + offP) + pc * 4;                                    // point boxes
    const char *pzs = reinterpret_cast<const char *>(sm + offP) + zsrc * 4;                                  // Z part region
    auto hbox = [&](int off, int c) { return reinterpret_cast<const float *>(hsc + off + (c < 2 ? c * HBOX_STRIDE : PT_S3OFF + (c - 2) * HBOX_STRIDE)); };
    auto xxbox = [&](int off) { return reinterpret_cast<const float *>(xsc + off); };
    auto lbox = [&](int off) { return reinterpret_cast<const LT *>(lsc + off); };
    // ring state as byte offsets / barrier addresses (no multiplies in the loop): halo slots of planes it, it+1,
    // it+2 (the one waited on inside the loop) and the point slot of plane it
    int ho = 0, ho1 = hstage, ho2 = 2 * hstage, po = 0;
    uint32_t hb0 = fullH, hb1 = fullH + 8, hb2 = fullH + 16, pbar = fullP;
    unsigned hpar = 0, ppar = 0;
    const int hend = nsh * hstage, pend = nsp * pstage;

    mbar_wait(fullH, 0);
    xx_p1 = xxbox(0)[0];
    if (sF[0] & TF_SHEAR) { xy_p1 = hbox(0, HB_SXY)[0]; xz_p1 = hbox(0, HB_SXZ)[0]; }
    mbar_wait(fullH + 8, 0);
    xx_p2 = xxbox(hstage)[0];

    for (int it = 0; it < nplanes; it++, q += s1, qy += qy_stride, qz += qz_stride) {
        const int i = ic0 + it;
        const unsigned f = sF[it];
        const bool fsh = f & TF_SHEAR;
        mbar_wait(hb2, hpar);
        xx_m1 = xx_0; xx_0 = xx_p1; xx_p1 = xx_p2; xx_p2 = xxbox(ho2)[0];
        xy_m2 = xy_m1; xy_m1 = xy_0; xy_0 = xy_p1;
        xz_m2 = xz_m1; xz_m1 = xz_0; xz_0 = xz_p1;
        if (sF[it + 1] & TF_SHEAR) { xy_p1 = hbox(ho1, HB_SXY)[0]; xz_p1 = hbox(ho1, HB_SXZ)[0]; }
        else { xy_p1 = 0.f; xz_p1 = 0.f; }
        mbar_wait(pbar, ppar);
        const bool xd = (f & TF_XD) != 0;
        const bool cellpml = xd || jkd;
        // the cell update, compiled twice: with and without the shear-stress differences (TF_SHEAR of this plane)
        auto cell_update = [&](auto shear_tag) {
            constexpr bool SHEAR = decltype(shear_tag)::value;
            const float *byy = hbox(ho, HB_SYY), *bzz = hbox(ho, HB_SZZ);
            const float *bxy = hbox(ho, HB_SXY), *bxz = hbox(ho, HB_SXZ), *byz = hbox(ho, HB_SYZ);
            const LT *l0p = lbox(ho), *l1p = lbox(ho1);
            const float *pb = reinterpret_cast<const float *>(psc + po);
            const unsigned l0 = l0p[0];
            const unsigned mi = l1p[0] & MSK, mj = l0p[LW] & MSK, mk = l0p[1] & MSK;
            float b0, bi, bj, bk;
            if (SMC) { b0 = sB[l0 & MSK]; bi = sB[mi]; bj = sB[mj]; bk = sB[mk]; }
            else { b0 = __ldg(&p.coef[l0 & MSK].B); bi = __ldg(&p.coef[mi].B); bj = __ldg(&p.coef[mj].B); bk = __ldg(&p.coef[mk].B); }
            const float bx = 0.5f * (b0 + bi), by = 0.5f * (b0 + bj), bz = 0.5f * (b0 + bk);
            float cib_a = BB_CA, cib_b = BB_CB, cif_a = BB_CA, cif_b = BB_CB;
            if (f & TF_IEDGE) { const AxisCoef ci = load_axis(p.axI, i); cib_a = ci.cab; cib_b = ci.cbb; cif_a = ci.caf; cif_b = ci.cbf; }
            float X[9];
            X[0] = D4C(cif_a, cif_b, xx_p1, xx_0, xx_p2, xx_m1);
            X[3] = D4C(cib_a, cib_b, xy_0, xy_m1, xy_p1, xy_m2);
            X[6] = D4C(cib_a, cib_b, xz_0, xz_m1, xz_p1, xz_m2);
            X[4] = D4C(cjf_a, cjf_b, byy[SW], byy[0], byy[2 * SW], byy[-SW]);
            X[8] = D4C(ckf_a, ckf_b, bzz[1], bzz[0], bzz[2], bzz[-1]);
            if constexpr (SHEAR) {
                X[1] = D4C(cjb_a, cjb_b, bxy[0], bxy[-SW], bxy[SW], bxy[-2 * SW]);
                X[2] = D4C(ckb_a, ckb_b, bxz[0], bxz[-1], bxz[1], bxz[-2]);
                X[5] = D4C(ckb_a, ckb_b, byz[0], byz[-1], byz[1], byz[-2]);
                X[7] = D4C(cjb_a, cjb_b, byz[0], byz[-SW], byz[SW], byz[-2 * SW]);
            } else { X[1] = X[2] = X[5] = X[7] = 0.f; }
            float v[3] = { pb[(QB_V + 0) * NT], pb[(QB_V + 1) * NT], pb[(QB_V + 2) * NT] };
            if (cellpml) {
                PmlCell pcell;
                pcell.xd = xd; pcell.jd = jd; pcell.kd = kd;
                const int ipx = i < p.P ? i - p.i0 : p.nxlo + (i - p.xhi_begin);
                pcell.qx = (unsigned)ipx * s1 + col; pcell.qy = qy; pcell.qz = qz;
                pcell.cI = p.axI + i; pcell.cJ = sJ + ty; pcell.cK = sK + tx;
                particle_pml<true>(p, pcell, bx, by, bz, X, v, pb + QB_X * NT, pb + (yoff >> 2), reinterpret_cast<const float *>(pzs + po), NT, zcomp);
            } else if constexpr (SHEAR) {
                v[0] += dt * bx * (X[0] + X[1] + X[2]);
                v[1] += dt * by * (X[3] + X[4] + X[5]);
                v[2] += dt * bz * (X[6] + X[7] + X[8]);
            } else {
                v[0] += dt * bx * X[0];
                v[1] += dt * by * (X[3] + X[4]);
                v[2] += dt * bz * (X[6] + X[8]);
            }
            if (l0 & LabelTraits<LT>::REFL) { v[0] = v[1] = v[2] = 0.f; }
            p.V[0][q] = v[0]; p.V[1][q] = v[1]; p.V[2][q] = v[2];
            if (pushsel) {
            float w0 = v[0], w1 = v[1], w2 = v[2];      // what the arrays hold after this half-step (the RMS maps below see v)
            if (p.bsrc_map) {      // particle sources of the boundary planes are injected here (the source kernel skips them)
                const int bp = boundary_plane(p, pushsel, i);
                float val, ox, oy, oz;
                if (bp >= 0 && boundary_source(p, bp, col, val, ox, oy, oz)) {
                    if (p.src_hard) { w0 = val * ox; w1 = val * oy; w2 = val * oz; } else { w0 += val * ox; w1 += val * oy; w2 += val * oz; }
                    p.V[0][q] = w0; p.V[1][q] = w1; p.V[2][q] = w2;
                }
            }
            if ((pushsel & 1) && i < p.i0 + 2) {
                float *b = p.peerV[0];
                const unsigned qn = (p.peer_plane[0] + (unsigned)(i - p.i0)) * s1 + col;
                b[qn] = w0; b[p.peer_vol[0] + qn] = w1; b[2 * p.peer_vol[0] + qn] = w2;
            }
            if ((pushsel & 2) && i >= p.i1 - 2) {
                float *b = p.peerV[1];
                const unsigned qn = (p.peer_plane[1] + (unsigned)(i - (p.i1 - 2))) * s1 + col;
                b[qn] = w0; b[p.peer_vol[1] + qn] = w1; b[2 * p.peer_vol[1] + qn] = w2;
            }
            }
            if (ACC && !cellpml) {
                const unsigned qa = q - 2 * s1;
                accumulate(p, BB_MAP_VX, qa, v[0], false);
                accumulate(p, BB_MAP_VY, qa, v[1], false);
                accumulate(p, BB_MAP_VZ, qa, v[2], false);
                accumulate(p, BB_MAP_ALLV, qa, v[0] * v[0] + v[1] * v[1] + v[2] * v[2], true);
            }
        };
#if BB_SKELETON == 0
        if (cellpml ? (upd_pml && !(f & TF_ILAST)) : active) {
            if (fsh) cell_update(std::true_type{}); else cell_update(std::false_type{});
        }
#elif BB_SKELETON == 2
        if (active) {
            const float *pb = reinterpret_cast<const float *>(psc + po);
            const float e = hbox(ho, HB_SYY)[0] + hbox(ho, HB_SZZ)[1] + xx_0;
            p.V[0][q] = pb[(QB_V + 0) * NT] + e; p.V[1][q] = pb[(QB_V + 1) * NT]; p.V[2][q] = pb[(QB_V + 2) * NT];
        }
#else
        (void)cell_update;
#endif
        __syncwarp();
        if (lane0) { mbar_arrive(hb0 + MAX_NSH * 8); mbar_arrive(pbar + MAX_NSP * 8); }
        ho = ho1; hb0 = hb1; ho1 = ho2; hb1 = hb2;
        ho2 += hstage; hb2 += 8;
        if (ho2 == hend) { ho2 = 0; hb2 = fullH; hpar ^= 1u; }
        po += pstage; pbar += 8;
        if (po == pend) { po = 0; pbar = fullP; ppar ^= 1u; }
    }
    if (p.dbg && tid == 0) {    // tid 0 is a consumer thread: it leaves the loop when the CTA's last plane is done
        unsigned long long *d = p.dbg + 4ull * ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x);
        d[0] = dbg_t0; d[1] = globaltimer_ns(); d[2] = ((unsigned long long)blockIdx.z << 40) | ((unsigned long long)blockIdx.y << 20) | blockIdx.x; d[3] = (unsigned long long)np;
    }
}
}  // namespace tma
