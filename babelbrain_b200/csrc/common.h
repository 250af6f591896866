// Shared host/device definitions for libbabelb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
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

// One row of the per-material coefficient table (32 B = one sector per lookup).
struct __align__(16) MatRow {
    float M, G, L, B;        // (lambda+2mu)/h, mu/h, lambda/h, 1/(rho h)   (relaxed moduli)
    float tauL, tauS, ots, K; // tau_L, tau_S, 1/tau_sigma, rho cL^2/h (pressure scaling)
};

enum { SP_VX_X = 0, SP_VX_Y, SP_VX_Z, SP_VY_X, SP_VY_Y, SP_VY_Z, SP_VZ_X, SP_VZ_Y, SP_VZ_Z,
       SP_SXX_X, SP_SXX_Y, SP_SXX_Z, SP_SYY_X, SP_SYY_Y, SP_SYY_Z, SP_SZZ_X, SP_SZZ_Y, SP_SZZ_Z,
       SP_SXY_X, SP_SXY_Y, SP_SXZ_X, SP_SXZ_Z, SP_SYZ_Y, SP_SYZ_Z, SP_COUNT };

// Device-side view of one slab.  Local plane ip <-> global i = i0 - 2 + ip (two halo planes on
// each side are always allocated); element (ip, j, k) lives at (ip*n2 + j)*pitch + k.
struct DevParams {
    int n1, n2, n3;      // global grid
    int i0, i1;          // owned planes
    int P;               // PML thickness
    int pitch;           // floats per k-row (multiple of 32)
    int nloc;            // i1 - i0 + 4
    long long plane;     // n2 * pitch
    float dt;
    float *V[3], *S[6], *R[6], *Pr;
    const void *lab;     // uint8_t or uint16_t labels, same layout; top bit = reflector
    const MatRow *mat;
    const float *pml;    // InvDXDT, DXDT, InvDXDThp, DXDThp, each P+1
    float *sp[SP_COUNT]; // split-field parts, compact over the PML shell of this slab
    int ilo_end, ihi_begin;
    long long off[6];
    // RMS / peak accumulators: [slot][(i-i0)*plane + j*pitch + k]
    float *acc_rms, *acc_peak;
    long long acc_stride;
    unsigned sel_maps;   // maps being accumulated
    int sel_rms_peak;
};
