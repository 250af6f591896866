/*
 * ORACLE (test infrastructure, NOT product code) -- plain C / OpenMP restatement of the
 * viscoelastic staggered-grid FDTD time loop that BabelBrain reaches through
 * PModel.StaggeredFDTD_3D_with_relaxation (TranscranialModeling/BabelIntegrationBASE.py:2338-2365).
 *
 * PARITY UNPINNED: the arithmetic lives in the un-vendored pip package BabelViscoFDTD
 * (==1.2.4, environment_linux.yml:44); no source, test or golden vector for it exists under
 * /root/reference and none can be fetched.  This restates the published scheme exactly as
 * oracle/fdtd_numpy.py does (same named items: differences with edge rules, tau-method memory
 * variables, split-field PML, soft sources, RMS/peak windows, sensor sampling) and is checked
 * against that float64 NumPy version in tests/test_oracle.py.
 *
 * Build: see oracle/Makefile (float32 -> liboracle_f32.so, -DORACLE_DOUBLE -> liboracle_f64.so).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it; it doubles as the reported OpenMP CPU baseline ("port").
 *
 * Layout: every volume is (N1,N2,N3) C-order, idx = (i*N2 + j)*N3 + k, as the caller's numpy
 * arrays are (BabelIntegrationBASE.py:2111, 2283).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef ORACLE_DOUBLE
typedef double real;
#define SQRT sqrt
#else
typedef float real;
#define SQRT sqrtf
#endif

#define CA ((real)(9.0 / 8.0))
#define CB ((real)(1.0 / 24.0))

enum { MAP_ALLV = 0, MAP_VX, MAP_VY, MAP_VZ, MAP_SXX, MAP_SYY, MAP_SZZ, MAP_SXY, MAP_SXZ, MAP_SYZ, MAP_P, MAP_COUNT };

typedef struct {
    int32_t n1, n2, n3, pml, nmat;
    int32_t nsrc, nt_src, steps, type_source;
    int32_t sel_rms_peak;       /* bit0 RMS, bit1 peak */
    uint32_t sel_maps_rms;      /* bit per MAP_* */
    uint32_t sel_maps_sensor;
    int32_t sensor_subsampling, sensor_start;
    int64_t nsrc_cells, nsensors;
    double dt;
    double mpml;                /* multi-axial damping ratio (oracle/fdtd_numpy.py: MPML_RATIO); 0 = classical layer */
} oracle_params;

typedef struct {
    int n1, n2, n3, P;
    int64_t s1, s2; /* strides of i and j */
    real dt;
    const uint32_t *mat;
    const real *M, *G, *L, *B, *tauL, *tauS, *ots, *K;
    const real *dmp, *dmphp;    /* damping at integer / half depth, P+1 each (oracle/fdtd_numpy.py: pml_damping) */
    real ratio;                 /* multi-axial damping ratio */
    real *V[3], *S[6], *R[6], *Pr;
    real *sp[24];
} ctx_t;

/* names of the 24 split arrays */
enum { VX_X = 0, VX_Y, VX_Z, VY_X, VY_Y, VY_Z, VZ_X, VZ_Y, VZ_Z,
       SXX_X, SXX_Y, SXX_Z, SYY_X, SYY_Y, SYY_Z, SZZ_X, SZZ_Y, SZZ_Z,
       SXY_X, SXY_Y, SXZ_X, SXZ_Z, SYZ_Y, SYZ_Z };

static inline int in_pml(int n, int N, int P) { return n < P || n >= N - P; }

/* integer-node and half-node damping along one axis (oracle/fdtd_numpy.py: pml_depth) */
static inline real damp_int(const ctx_t *c, int n, int N) {
    int d = 0;
    if (n < c->P) d = c->P - n; else if (n >= N - c->P) d = n - (N - c->P - 1);
    return c->dmp[d];
}
static inline real damp_half(const ctx_t *c, int n, int N) {
    if (n < c->P) return c->dmphp[c->P - 1 - n];
    if (n >= N - c->P) return c->dmphp[n - (N - c->P - 1)];
    return 0;
}
/* (InvDXDT, DXDT) of a split part: own damping plus ratio x the integer-node damping of the two other axes
   (oracle/fdtd_numpy.py: coef) */
static inline void coef_of(const ctx_t *c, real own, real o1, real o2, real *a, real *b) {
    const real d = own + c->ratio * (o1 + o2);
    *a = 1 / (1 / c->dt + d / 2); *b = 1 / c->dt - d / 2;
}

/* backward / forward staggered differences with the edge rules; st = stride of the axis */
static inline real dbwd(const real *f, int64_t p, int64_t st, int n, int N) {
    if (n > 1 && n < N - 1) return CA * (f[p] - f[p - st]) - CB * (f[p + st] - f[p - 2 * st]);
    if (n > 0) return f[p] - f[p - st];
    return 0;
}
static inline real dfwd(const real *f, int64_t p, int64_t st, int n, int N) {
    if (n > 0 && n < N - 2) return CA * (f[p + st] - f[p]) - CB * (f[p + 2 * st] - f[p - st]);
    if (n < N - 1) return f[p + st] - f[p];
    return 0;
}

static inline real harm4(real g1, real g2, real g3, real g4) {
    if (g1 * g2 * g3 * g4 == 0) return 0;
    return (real)4 / ((real)1 / g1 + (real)1 / g2 + (real)1 / g3 + (real)1 / g4);
}

static inline real splitupd(real *arr, int64_t p, real a, real b, real C, real D) {
    real v = a * (arr[p] * b + C * D);
    arr[p] = v;
    return v;
}

static void stress_cell(const ctx_t *c, int i, int j, int k, real *out_p_acc) {
    const int n1 = c->n1, n2 = c->n2, n3 = c->n3;
    const int64_t s1 = c->s1, s2 = c->s2, p = i * s1 + j * s2 + k;
    const real dt = c->dt;
    const real *Vx = c->V[0], *Vy = c->V[1], *Vz = c->V[2];
    const uint32_t m = c->mat[p];
    const int pml = in_pml(i, n1, c->P) || in_pml(j, n2, c->P) || in_pml(k, n3, c->P);
    if (pml && !(i < n1 - 1 && j < n2 - 1 && k < n3 - 1)) return;
    /* neighbour labels for the edge (shear) quantities */
    const uint32_t mi = c->mat[p + s1], mj = c->mat[p + s2], mk = c->mat[p + 1];
    const uint32_t mij = c->mat[p + s1 + s2], mik = c->mat[p + s1 + 1], mjk = c->mat[p + s2 + 1];
    const real rigxy = harm4(c->G[m], c->G[mi], c->G[mj], c->G[mij]);
    const real rigxz = harm4(c->G[m], c->G[mi], c->G[mk], c->G[mik]);
    const real rigyz = harm4(c->G[m], c->G[mj], c->G[mk], c->G[mjk]);
    const real Dxx = dbwd(Vx, p, s1, i, n1), Dyy = dbwd(Vy, p, s2, j, n2), Dzz = dbwd(Vz, p, 1, k, n3);
    if (pml) {
        real a, b;
        const real M = c->M[m], L = c->L[m];
        const real di = damp_int(c, i, n1), dj = damp_int(c, j, n2), dk = damp_int(c, k, n3);
        coef_of(c, di, dj, dk, &a, &b);
        real xx = splitupd(c->sp[SXX_X], p, a, b, M, Dxx);
        real yy = splitupd(c->sp[SYY_X], p, a, b, L, Dxx);
        real zz = splitupd(c->sp[SZZ_X], p, a, b, L, Dxx);
        coef_of(c, dj, di, dk, &a, &b);
        xx += splitupd(c->sp[SXX_Y], p, a, b, L, Dyy);
        yy += splitupd(c->sp[SYY_Y], p, a, b, M, Dyy);
        zz += splitupd(c->sp[SZZ_Y], p, a, b, L, Dyy);
        coef_of(c, dk, di, dj, &a, &b);
        xx += splitupd(c->sp[SXX_Z], p, a, b, L, Dzz);
        yy += splitupd(c->sp[SYY_Z], p, a, b, L, Dzz);
        zz += splitupd(c->sp[SZZ_Z], p, a, b, M, Dzz);
        c->S[0][p] = xx; c->S[1][p] = yy; c->S[2][p] = zz;
        real ai, bi, aj, bj, ak, bk;
        coef_of(c, damp_half(c, i, n1), dj, dk, &ai, &bi);
        coef_of(c, damp_half(c, j, n2), di, dk, &aj, &bj);
        coef_of(c, damp_half(c, k, n3), di, dj, &ak, &bk);
        c->S[3][p] = splitupd(c->sp[SXY_X], p, ai, bi, rigxy, dfwd(Vy, p, s1, i, n1))
                   + splitupd(c->sp[SXY_Y], p, aj, bj, rigxy, dfwd(Vx, p, s2, j, n2));
        c->S[4][p] = splitupd(c->sp[SXZ_X], p, ai, bi, rigxz, dfwd(Vz, p, s1, i, n1))
                   + splitupd(c->sp[SXZ_Z], p, ak, bk, rigxz, dfwd(Vx, p, 1, k, n3));
        c->S[5][p] = splitupd(c->sp[SYZ_Y], p, aj, bj, rigyz, dfwd(Vz, p, s2, j, n2))
                   + splitupd(c->sp[SYZ_Z], p, ak, bk, rigyz, dfwd(Vy, p, 1, k, n3));
        return;
    }
    /* interior */
    const real th = Dxx + Dyy + Dzz;
    const real tL = c->tauL[m], tS = c->tauS[m], ots = c->ots[m];
    const real LM = c->M[m] * (1 + tL), Mi2 = 2 * c->G[m] * (1 + tS);
    const int att = (tL != 0) || (tS != 0);
    const real LMC = dt * c->M[m] * (tL * ots), MC = dt * 2 * c->G[m] * (tS * ots);
    const real den = 1 + dt * (real)0.5 * ots, num = 1 - dt * (real)0.5 * ots;
    c->Pr[p] += dt * th;
    const real oth[3] = { Dyy + Dzz, Dxx + Dzz, Dxx + Dyy };
    for (int q = 0; q < 3; q++) {
        if (att) {
            const real R = c->R[q][p];
            const real NextR = (num * R - LMC * th + MC * oth[q]) / den;
            c->S[q][p] += dt * (LM * th - Mi2 * oth[q] + (real)0.5 * (R + NextR));
            c->R[q][p] = NextR;
        } else {
            c->S[q][p] += dt * (LM * th - Mi2 * oth[q]);
        }
    }
    const real rig[3] = { rigxy, rigxz, rigyz };
    const real tsum[3] = { c->tauS[m] + c->tauS[mi] + c->tauS[mj] + c->tauS[mij],
                           c->tauS[m] + c->tauS[mi] + c->tauS[mk] + c->tauS[mik],
                           c->tauS[m] + c->tauS[mj] + c->tauS[mk] + c->tauS[mjk] };
    for (int q = 0; q < 3; q++) {
        if (rig[q] == 0) continue;
        real D;
        if (q == 0) D = dfwd(Vy, p, s1, i, n1) + dfwd(Vx, p, s2, j, n2);
        else if (q == 1) D = dfwd(Vz, p, s1, i, n1) + dfwd(Vx, p, 1, k, n3);
        else D = dfwd(Vz, p, s2, j, n2) + dfwd(Vy, p, 1, k, n3);
        const real te = (real)0.25 * tsum[q];
        if (te != 0) {
            const real R = c->R[3 + q][p];
            const real NextR = (num * R - dt * (rig[q] * (te * ots)) * D) / den;
            c->S[3 + q][p] += dt * (rig[q] * (1 + te) * D + (real)0.5 * (R + NextR));
            c->R[3 + q][p] = NextR;
        } else {
            c->S[3 + q][p] += dt * (rig[q] * (1 + te) * D);
        }
    }
    (void)out_p_acc;
}

static void particle_cell(const ctx_t *c, int i, int j, int k) {
    const int n1 = c->n1, n2 = c->n2, n3 = c->n3;
    const int64_t s1 = c->s1, s2 = c->s2, p = i * s1 + j * s2 + k;
    const real dt = c->dt;
    const int pml = in_pml(i, n1, c->P) || in_pml(j, n2, c->P) || in_pml(k, n3, c->P);
    if (pml && !(i < n1 - 1 && j < n2 - 1 && k < n3 - 1)) return;
    const real *Sxx = c->S[0], *Syy = c->S[1], *Szz = c->S[2], *Sxy = c->S[3], *Sxz = c->S[4], *Syz = c->S[5];
    const real b0 = c->B[c->mat[p]];
    const real bx = (real)0.5 * (b0 + c->B[c->mat[p + s1]]);
    const real by = (real)0.5 * (b0 + c->B[c->mat[p + s2]]);
    const real bz = (real)0.5 * (b0 + c->B[c->mat[p + 1]]);
    const real x1 = dfwd(Sxx, p, s1, i, n1), x2 = dbwd(Sxy, p, s2, j, n2), x3 = dbwd(Sxz, p, 1, k, n3);
    const real y1 = dbwd(Sxy, p, s1, i, n1), y2 = dfwd(Syy, p, s2, j, n2), y3 = dbwd(Syz, p, 1, k, n3);
    const real z1 = dbwd(Sxz, p, s1, i, n1), z2 = dbwd(Syz, p, s2, j, n2), z3 = dfwd(Szz, p, 1, k, n3);
    if (pml) {
        real ai, bi, aj, bj, ak, bk, hi, gi, hj, gj, hk, gk;
        const real di = damp_int(c, i, n1), dj = damp_int(c, j, n2), dk = damp_int(c, k, n3);
        coef_of(c, di, dj, dk, &ai, &bi); coef_of(c, dj, di, dk, &aj, &bj); coef_of(c, dk, di, dj, &ak, &bk);
        coef_of(c, damp_half(c, i, n1), dj, dk, &hi, &gi);
        coef_of(c, damp_half(c, j, n2), di, dk, &hj, &gj);
        coef_of(c, damp_half(c, k, n3), di, dj, &hk, &gk);
        c->V[0][p] = splitupd(c->sp[VX_X], p, hi, gi, bx, x1) + splitupd(c->sp[VX_Y], p, aj, bj, bx, x2)
                   + splitupd(c->sp[VX_Z], p, ak, bk, bx, x3);
        c->V[1][p] = splitupd(c->sp[VY_X], p, ai, bi, by, y1) + splitupd(c->sp[VY_Y], p, hj, gj, by, y2)
                   + splitupd(c->sp[VY_Z], p, ak, bk, by, y3);
        c->V[2][p] = splitupd(c->sp[VZ_X], p, ai, bi, bz, z1) + splitupd(c->sp[VZ_Y], p, aj, bj, bz, z2)
                   + splitupd(c->sp[VZ_Z], p, hk, gk, bz, z3);
        return;
    }
    c->V[0][p] += dt * bx * (x1 + x2 + x3);
    c->V[1][p] += dt * by * (y1 + y2 + y3);
    c->V[2][p] += dt * bz * (z1 + z2 + z3);
}

static inline real map_value(const ctx_t *c, int map, int64_t p) {
    switch (map) {
    case MAP_ALLV: return c->V[0][p] * c->V[0][p] + c->V[1][p] * c->V[1][p] + c->V[2][p] * c->V[2][p];
    case MAP_VX: case MAP_VY: case MAP_VZ: return c->V[map - MAP_VX][p];
    case MAP_P: return -c->K[c->mat[p]] * c->Pr[p];
    default: return c->S[map - MAP_SXX][p];
    }
}

static int popcount_below(uint32_t mask, int bit) { return __builtin_popcount(mask & ((1u << bit) - 1)); }

/*
 * tables: 8 arrays of nmat (M,G,L,B,tauL,tauS,ots,K); pmltab: 2 arrays of P+1 (damping at integer / half depth);
 * src_cell: C-order linear index of each source cell; src_id 0-based row; o_xyz per source cell;
 * srcfun [nt_src][nsrc]; sensor_cell C-order linear index in the order of IndexSensorMap.
 * out_rms / out_peak: [selected map (ascending MAP_* bit)][N]; out_sensor: [selected map][nsensors][nsamples];
 * out_last: [10 (Vx..Syz, P)][N] or NULL.
 */
int oracle_fdtd_run(const oracle_params *prm, const uint32_t *matmap, const real *tables, const real *pmltab,
                    const int64_t *src_cell, const int32_t *src_id, const real *ox, const real *oy,
                    const real *oz, const real *srcfun, const int64_t *sensor_cell,
                    const uint32_t *reflector, real *out_rms, real *out_peak, real *out_sensor,
                    real *out_last) {
    ctx_t c;
    memset(&c, 0, sizeof(c));
    c.n1 = prm->n1; c.n2 = prm->n2; c.n3 = prm->n3; c.P = prm->pml;
    c.s2 = c.n3; c.s1 = (int64_t)c.n2 * c.n3;
    c.dt = (real)prm->dt;
    c.mat = matmap;
    const int nm = prm->nmat, P1 = prm->pml + 1;
    c.M = tables; c.G = tables + nm; c.L = tables + 2 * nm; c.B = tables + 3 * nm;
    c.tauL = tables + 4 * nm; c.tauS = tables + 5 * nm; c.ots = tables + 6 * nm; c.K = tables + 7 * nm;
    c.dmp = pmltab; c.dmphp = pmltab + P1; c.ratio = (real)prm->mpml;
    const int64_t N = (int64_t)c.n1 * c.n2 * c.n3;
    const int narr = 3 + 6 + 6 + 1 + 24;
    real *pool = (real *)calloc((size_t)narr * N, sizeof(real));
    if (!pool) return -1;
    real *q = pool;
    for (int a = 0; a < 3; a++) { c.V[a] = q; q += N; }
    for (int a = 0; a < 6; a++) { c.S[a] = q; q += N; }
    for (int a = 0; a < 6; a++) { c.R[a] = q; q += N; }
    c.Pr = q; q += N;
    for (int a = 0; a < 24; a++) { c.sp[a] = q; q += N; }

    const int sub = prm->sensor_subsampling;
    const int n0 = prm->sensor_start * sub;
    int nsamples = 0;
    for (int n = 0; n < prm->steps; n++) if (n % sub == 0 && n / sub >= prm->sensor_start) nsamples++;
    const int nsel = __builtin_popcount(prm->sel_maps_rms);
    if ((prm->sel_rms_peak & 1) && out_rms) memset(out_rms, 0, sizeof(real) * nsel * N);
    if ((prm->sel_rms_peak & 2) && out_peak) memset(out_peak, 0, sizeof(real) * nsel * N);
    const int n1 = c.n1, n2 = c.n2, n3 = c.n3, P = c.P;
    const uint32_t stress_maps = prm->sel_maps_rms & 0x7F0u, part_maps = prm->sel_maps_rms & 0xFu;
    int sample = 0;
    for (int n = 0; n < prm->steps; n++) {
#pragma omp parallel for collapse(2) schedule(static)
        for (int i = 0; i < n1; i++)
            for (int j = 0; j < n2; j++)
                for (int k = 0; k < n3; k++) stress_cell(&c, i, j, k, 0);
        for (int pass = 0; pass < 2; pass++) {
            /* pass 0: stress half-step already done above; pass 1: particle half-step.
               Order inside a half-step: update -> reflector -> RMS/peak -> source. */
            if (pass == 1) {
#pragma omp parallel for collapse(2) schedule(static)
                for (int i = 0; i < n1; i++)
                    for (int j = 0; j < n2; j++)
                        for (int k = 0; k < n3; k++) particle_cell(&c, i, j, k);
            }
            if (reflector) {
#pragma omp parallel for
                for (int64_t p = 0; p < N; p++) {
                    if (!reflector[p]) continue;
                    if (pass == 0) { for (int a = 0; a < 6; a++) c.S[a][p] = 0; c.Pr[p] = 0; }
                    else for (int a = 0; a < 3; a++) c.V[a][p] = 0;
                }
            }
            const uint32_t maps = pass == 0 ? stress_maps : part_maps;
            if (n >= n0 && maps) {
#pragma omp parallel for collapse(2) schedule(static)
                for (int i = P; i < n1 - P; i++)
                    for (int j = P; j < n2 - P; j++)
                        for (int k = P; k < n3 - P; k++) {
                            const int64_t p = i * c.s1 + j * c.s2 + k;
                            for (int m = 0; m < MAP_COUNT; m++) {
                                if (!(maps & (1u << m))) continue;
                                const real v = map_value(&c, m, p);
                                const int64_t o = (int64_t)popcount_below(prm->sel_maps_rms, m) * N + p;
                                if (prm->sel_rms_peak & 1) out_rms[o] += (m == MAP_ALLV) ? v : v * v;
                                if ((prm->sel_rms_peak & 2) && v > out_peak[o]) out_peak[o] = v;
                            }
                        }
            }
            if (n < prm->nt_src) {
                for (int64_t s = 0; s < prm->nsrc_cells; s++) {
                    const real v = srcfun[(int64_t)n * prm->nsrc + src_id[s]];
                    const int64_t p = src_cell[s];
                    if (pass == 0 && prm->type_source >= 2) {
                        const real w = v * ox[s];
                        if (prm->type_source == 2) { c.S[0][p] += w; c.S[1][p] += w; c.S[2][p] += w; }
                        else { c.S[0][p] = w; c.S[1][p] = w; c.S[2][p] = w; }
                    } else if (pass == 1 && prm->type_source < 2) {
                        if (prm->type_source == 0) { c.V[0][p] += v * ox[s]; c.V[1][p] += v * oy[s]; c.V[2][p] += v * oz[s]; }
                        else { c.V[0][p] = v * ox[s]; c.V[1][p] = v * oy[s]; c.V[2][p] = v * oz[s]; }
                    }
                }
            }
        }
        if (n % sub == 0 && n / sub >= prm->sensor_start) {
            for (int m = 0; m < MAP_COUNT; m++) {
                if (!(prm->sel_maps_sensor & (1u << m))) continue;
                real *o = out_sensor + (int64_t)popcount_below(prm->sel_maps_sensor, m) * prm->nsensors * nsamples;
#pragma omp parallel for
                for (int64_t s = 0; s < prm->nsensors; s++) {
                    real v = map_value(&c, m, sensor_cell[s]);
                    if (m == MAP_ALLV) v = SQRT(v);
                    o[s * nsamples + sample] = v;
                }
            }
            sample++;
        }
    }
    const real nacc = (real)(prm->steps - n0 > 0 ? prm->steps - n0 : 1);
    if (prm->sel_rms_peak & 1)
        for (int64_t p = 0; p < (int64_t)nsel * N; p++) out_rms[p] = SQRT(out_rms[p] / nacc);
    if ((prm->sel_rms_peak & 2) && (prm->sel_maps_rms & 1))
        for (int64_t p = 0; p < N; p++) out_peak[p] = SQRT(out_peak[p]);
    if (out_last) {
        for (int m = MAP_VX; m <= MAP_P; m++)
            for (int64_t p = 0; p < N; p++) out_last[(int64_t)(m - 1) * N + p] = map_value(&c, m, p);
    }
    free(pool);
    return nsamples;
}

int oracle_real_size(void) { return (int)sizeof(real); }
int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
/* torch.distributed.run exports OMP_NUM_THREADS=1 to its children; the CPU baseline asks for the host's cores explicitly */
void oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
