/*
 * ORACLE (test infrastructure, NOT product code) -- plain C restatement of the Rayleigh-Sommerfeld
 * source-field integral BabelBrain calls as ForwardSimple(cwvnb, center, ds, u0, rf)
 * (e.g. TranscranialModeling/BabelIntegrationSingle.py:295, BabelIntegrationANNULAR_ARRAY.py:411).
 *
 * PARITY UNPINNED: ForwardSimple lives in BabelViscoFDTD.tools.RayleighAndBHTE (un-vendored pip
 * package, ==1.2.4 environment_linux.yml:44).  Restated from the published integral:
 *   S_p  = sum_s ds_s * exp(Im(k) R) / R * u0_s * exp(-j Re(k) R),   R = |rf_p - center_s|
 *   out_p = j k S_p / (2 pi)
 * with an optional MaxDistance skip (never passed by BabelBrain).  R == 0 yields inf/NaN locally.
 */
#include <stdint.h>
#include <math.h>
#ifdef ORACLE_DOUBLE
typedef double real;
#define SQRT sqrt
#define EXP exp
#define SIN sin
#define COS cos
#else
typedef float real;
#define SQRT sqrtf
#define EXP expf
#define SIN sinf
#define COS cosf
#endif

void oracle_rayleigh_forward(real k_re, real k_im, int64_t nsrc, const real *center, const real *ds,
                             const real *u0_reim, int64_t npts, const real *rf, real *out_reim,
                             real max_distance) {
    const real two_pi = (real)6.283185307179586476925286766559;
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < npts; p++) {
        const real x = rf[3 * p], y = rf[3 * p + 1], z = rf[3 * p + 2];
        real sr = 0, si = 0;
        for (int64_t s = 0; s < nsrc; s++) {
            const real dx = center[3 * s] - x, dy = center[3 * s + 1] - y, dz = center[3 * s + 2] - z;
            const real R = SQRT(dx * dx + dy * dy + dz * dz);
            if (max_distance > 0 && R > max_distance) continue;
            const real amp = EXP(R * k_im) * ds[s] / R;
            const real cs = COS(R * k_re), sn = SIN(R * k_re);
            const real ur = u0_reim[2 * s], ui = u0_reim[2 * s + 1];
            sr += amp * (ur * cs + ui * sn);
            si += amp * (ui * cs - ur * sn);
        }
        out_reim[2 * p] = (-sr * k_im - si * k_re) / two_pi;
        out_reim[2 * p + 1] = (sr * k_re - si * k_im) / two_pi;
    }
}
