// Shared host/device definitions for libbabelb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>      // CUtensorMap (types only: the driver entry point is resolved at run time)
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/babelb200.h"

void bb_set_error(const char *fmt, ...);

#define BB_CUDA(call)                                                                       \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            bb_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,               \
                         cudaGetErrorString(e__));                                          \
            return BB_ERR_CUDA;                                                             \
        }                                                                                   \
    } while (0)

#define BB_REQUIRE(cond, ...)                                                               \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            bb_set_error(__VA_ARGS__);                                                      \
            return BB_ERR_ARG;                                                              \
        }                                                                                   \
    } while (0)

// Derived per-material coefficients (computed once, in double, from the host's BB_NCOEF-float row
// and dt) so that the per-cell update has no divisions:
//   normal:  R' = a R - LMCb th + MCb oth ;  S += dt (LM th - Mi2 oth + (R+R')/2)
//   edge  :  rig = 4/(sum invG) (0 when any neighbour is fluid: invG = +inf), te = mean tauS,
//            R' = a R - cs rig te D ;  S += dt (rig (1+te) D + (R+R')/2)
struct __align__(16) MatCoef {
    float LM, Mi2, LMCb, MCb;   // M(1+tauL), 2G(1+tauS), dt M tauL ots/den, dt 2G tauS ots/den
    float a, cs, K, invG;       // num/den, dt ots/den, rho cL^2/h, 1/G (+inf for fluids)
    float tauS, B, M, L;        // tau_S, 1/(rho h), relaxed (lambda+2mu)/h and lambda/h (PML)
};
constexpr int BB_MAX_SMEM_MAT = 128;   // uint8 labels hold at most 127 materials -> table in smem

// Per-index coefficients of one grid axis (built on the host from the PML table):
//   eI, eH = half the damping d/2 of the axis at the integer node n and at the half node n + 1/2 (0 outside the PML
//   of that axis).  A split part advances as f' = a f + b C D with b = 1/(1/dt + e), a = (1/dt - e) b, where
//   e = e_own(staggering of the part) + mpml * (eI of the two other axes): the multi-axial layer of Meza-Fajardo &
//   Papageorgiou (2008).  The classical layer (mpml = 0) is unstable where a fluid-solid interface enters the shell
//   (DESIGN.md section 4.3), so the coefficients are formed per cell from the three axes' values;
//   staggered differences  D- = cab (f0 - f-1) - cbb (f+1 - f-2),  D+ = caf (f+1 - f0) - cbf (f+2 - f-1)
//   with the domain-edge rules folded into the coefficients (9/8,1/24 | 1,0 | 0,0).
struct __align__(16) AxisCoef {
    float eI, eH, pad0, pad1;
    float cab, cbb, caf, cbf;
};

// split-field parts kept per damping axis (8 fields each), only where that axis is damped:
//   X parts: Sxx Syy Szz Sxy Sxz | Vx Vy Vz     (planes with i in the PML)
//   Y parts: Sxx Syy Szz Sxy Syz | Vx Vy Vz     (rows with j in the PML)
//   Z parts: Sxx Syy Szz Sxz Syz | Vx Vy Vz     (columns with k in the PML)
constexpr int BB_NPART = 8;

// rows of a (j,k) tile of the half-step kernels: the Y parts are stored for whole tile rows
#ifndef BB_TILE_ROWS
#define BB_TILE_ROWS 8
#endif
constexpr int BB_TY = BB_TILE_ROWS;

// tile flags, one byte per (local plane, tile row, tile column) of the 8x64 (j,k) tiling
enum { TF_ATT = 1,      // a non-PML cell of the tile attenuates -> normal memory variables move
       TF_SOLID = 2,    // a cell of the tile (+1 in j,k, planes i and i+1) has G != 0 -> shear stresses / memory variables move
       TF_SHEAR = 4,    // a cell of the tile (+-2 in j,k) on this plane has G != 0 -> its shear stresses can be non-zero
       TF_INT = 8,      // the tile holds non-PML cells on this plane
       // properties of the plane itself (the same for every tile), so that the plane loop tests bits instead of comparing indices:
       TF_XD = 16,      // the plane lies inside the i-PML
       TF_IEDGE = 32,   // i <= 1 or i >= n1-2: the i-differences use the domain-edge coefficients
       TF_ILAST = 64 }; // i == n1-1 (split-field cells of the last plane are not updated)

// Device-side view of one slab.  Local plane ip <-> global i = i0 - 2 + ip (two halo planes on
// each side are always allocated); element (ip, j, k) lives at (ip*n2 + j)*pitch + k.
struct DevParams {
    int n1, n2, n3;      // global grid
    int i0, i1;          // owned planes
    int P;               // PML thickness
    int pitch;           // floats per k-row (multiple of 32)
    int nloc;            // i1 - i0 + 4
    long long plane;     // n2 * pitch
    float dt, idt;       // time step and 1/dt
    float mpml;          // multi-axial damping ratio of the PML (bb_fdtd_desc::mpml_ratio)
    float *V[3], *S[6], *R[6], *Pr;
    const void *lab;     // uint8_t or uint16_t labels, same layout; top bit = reflector
    const MatCoef *coef; // derived rows indexed by label
    int nmat;
    const unsigned char *flags; // [nloc][ntj][ntk]
    int ntj, ntk;
    const AxisCoef *axI, *axJ, *axK;   // [n1], [n2], [n3]
    // damped parts, stored tile-aligned so that a TMA box of a part maps 1:1 onto a tile:
    //   XP [(ipx*n2 + j)*pitch + k]              over this slab's i-PML planes,
    //   YP [((i-i0)*nyrows + jp)*pitch + k]      jp = ytile(j/8)*8 + j%8 over the tile rows that hold j-PML cells,
    //   ZP [((i-i0)*n2 + j)*zpw + kp]            kp = k (low side) or zbw + k - (n3-P) (high side); zbw = P rounded up to 4
    // with ytile(t) = t < nylo ? t : t - tjhi0 + nylo.
    float *XP[BB_NPART], *YP[BB_NPART], *ZP[BB_NPART];
    int nxlo;            // owned planes inside the low-i PML: global i in [i0, i0+nxlo)
    int xhi_begin;       // first owned plane inside the high-i PML (== i1 when none)
    int nylo, tjhi0, nyrows;   // j tiles [0,nylo) and [tjhi0,ntj) hold PML rows; nyrows = stored rows
    int zbw, zpw;              // columns stored per side of the k-PML, and per row (2 zbw)
    // slab neighbours reached through NVLink peer memory (index 0: lower neighbour, 1: upper).  The boundary CTAs of
    // a half-step store the two planes the neighbour needs straight into its halo planes and, when the last of
    // them is done, publish the half-step's sequence number in the neighbour's flag word.
    float *peerV[2], *peerS[2];          // base of the neighbour's V / S component groups (nullptr: no such neighbour)
    unsigned peer_plane[2];              // neighbour-local plane that receives this slab's first pushed plane of that side
    long long peer_vol[2];               // component pitch of the neighbour's groups (its nloc * plane)
    unsigned long long *flag_local;      // [2] written by the lower / upper neighbour
    unsigned long long *flag_peer[2];    // where this slab publishes: lower neighbour's [1], upper neighbour's [0]
    unsigned *push_count;                // [2] plane-pushes completed so far in this launch (last CTA publishes)
    unsigned long long seq;              // (epoch << 32) | (half-step index + 1) of this launch
    int publish;                         // 1: the half-step kernel publishes seq itself (always, since the boundary-plane sources moved into it)
    int exp;                             // timing experiments only (BB_EXPERIMENT_HALO): 1 no fence before the count, 4 no flag wait, 16 system-scope fence after the flag wait, 32 system-scope fence before every count
    unsigned long long peer_timeout_ns;  // longest wait for a neighbour's halo (0: forever); on a timeout *err is set and the run fails
    int *err;                            // device error word checked by bb_fdtd_run (1, 2: halo wait on the lower / upper side timed out)
    // Sources that sit in the slab's boundary planes (the two planes next to each existing neighbour) are injected by the
    // half-step kernel itself, so that those planes can be pushed and published without waiting for the source kernel
    // that follows: bsrc_map[b * plane + j * pitch + k] = index into the (boundary-first) source-cell arrays or -1, for
    // b = 0, 1 (planes i0, i0+1) and 2, 3 (planes i1-2, i1-1).  nullptr when the launch injects no sources of its kind.
    const int *bsrc_map;
    const int *bsrc_row;                 // source row of each source cell
    const float *bsrc_o[3];              // Ox / Oy / Oz weights of each source cell
    const float *sf_row;                 // SourceFunctions row of this time step (nullptr: continuous-wave tones)
    const float *tone_ac, *tone_as;      // per source A cos(phi), A sin(phi)
    float env_sin, env_cos;              // ramp(n) sin(w t_n), ramp(n) cos(w t_n)
    int src_hard;                        // 0 soft (additive), 1 hard (replaces the field)
    // profiling aid (BB_CTA_TIMING=1): per CTA of the last launch {globaltimer at start, at end, blockIdx packed, planes}
    unsigned long long *dbg;
    // RMS / peak accumulators: [slot][(i-i0)*plane + j*pitch + k]
    float *acc_rms, *acc_peak;
    long long acc_stride;
    unsigned sel_maps;   // maps being accumulated
    int sel_rms_peak;
};

// Plane ranges of the CTAs of one launch along the marching axis: chunk z covers planes [start[z], end[z]).
// Long chunks first, short ones last, so the last wave of CTAs is short (the hardware hands out CTAs in order).
constexpr int BB_MAX_CHUNKS = 48;
struct ChunkPlan { int n; int start[BB_MAX_CHUNKS], end[BB_MAX_CHUNKS]; };

// TMA descriptors (passed as a __grid_constant__ parameter).  The component arrays of a field group
// (V[3], S[6], R[6], parts[8]) are contiguous, so the group is a 4-D tensor (k, j, plane, component)
// and one TMA instruction moves the boxes of 2 or 3 components at once.
struct StressMaps {
    CUtensorMap v3;      // V, box (TX+8, TY+4, 1, 3)
    CUtensorMap lab;     // labels, box (LW, TY+1, 1)
    CUtensorMap s3;      // S, box (TX, TY, 1, 3): component 0 = normal, 3 = shear stresses
    CUtensorMap r3;      // R, same
    CUtensorMap pr;      // pressure accumulator, box (TX, TY, 1)
    CUtensorMap xp3, xp2, yp3, yp2, zp3, zp2;   // damped parts: box depth 3 (normal, component 0) / 2 (shear, component 3)
    CUtensorMap acc;     // pressure RMS accumulator (slot 0), box (TX, TY, 1)
};
struct ParticleMaps {
    CUtensorMap sxx;     // S component 0, box (TX, TY, 1, 1): only the i-stencil
    CUtensorMap sh2;     // S components 1-2 (Syy Szz), box (TX+8, TY+4, 1, 2)
    CUtensorMap sh3;     // S components 3-5 (Sxy Sxz Syz), box (TX+8, TY+4, 1, 3)
    CUtensorMap lab;
    CUtensorMap v3;      // V, box (TX, TY, 1, 3)
    CUtensorMap xp3, yp3, zp3;   // damped velocity parts: components 5-7 of the part groups
};
