/*
 * babelb200.h -- C ABI of libbabelb200.so: the B200 (sm_100a) replacement for the solver calls
 * BabelBrain makes into the BabelViscoFDTD package.
 *
 * The reference binds this path from Python, so the binding a maintainer adds is a ctypes stub
 * (babelbrain_b200/_capi.py; INTEGRATION.md shows it).  Every entry point below replaces one
 * piece of a reference call:
 *
 *   bb_fdtd_*            PModel.StaggeredFDTD_3D_with_relaxation(...)
 *                        TranscranialModeling/BabelIntegrationBASE.py:2338-2365 (forward),
 *                        :2374-2398 (back-propagation), :2401-2428 (refocus)
 *   bb_rayleigh_forward  ForwardSimple(cwvnb, center, ds, u0, rf)
 *                        TranscranialModeling/BabelIntegrationSingle.py:295,
 *                        BabelIntegrationANNULAR_ARRAY.py:383,411,
 *                        BabelIntegrationCONCAVE_PHASEDARRAY.py:307,328,425,446
 *   bb_bhte_run          BHTE / BHTEMultiplePressureFields(...)
 *                        ThermalModeling/CalculateTemperatureEffects.py:365-395, :406, :439, :960-990
 *   bb_device_count/name InitCuda(deviceName) (BabelIntegrationBASE.py:918-925) and
 *                        StaggeredFDTD_3D_With_Relaxation_CUDA.ListDevices()
 *                        (BabelBrain/SelFiles/SelFiles.py:254-258)
 *
 * Conventions: plain pointers and sizes only; the caller owns every host buffer, the library
 * owns all device memory inside the opaque handle; every function returns 0 on success and a
 * non-zero code otherwise, with a thread-local message in bb_last_error().  There is no CPU
 * fallback: without a CUDA device every compute entry point fails with BB_ERR_CUDA.
 *
 * Volumes are (N1,N2,N3) C-order (k fastest), exactly the caller's numpy arrays
 * (BabelIntegrationBASE.py:2111 MaterialMap uint32, :2283 SensorMap uint32).  A handle owns the
 * slab of planes i in [i0,i1) of the global grid (i0=0,i1=N1 on one GPU); host volume pointers
 * passed to a handle cover the planes [max(i0-2,0), min(i1+2,N1)) -- the slab plus its halo.
 */
#ifndef BABELB200_H
#define BABELB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BB_OK 0
#define BB_ERR_ARG 1
#define BB_ERR_CUDA 2
#define BB_ERR_NCCL 3
#define BB_ERR_STATE 4

/* map ids = bit positions of the SelMaps masks (order of the reference's map names) */
enum bb_map {
    BB_MAP_ALLV = 0, BB_MAP_VX, BB_MAP_VY, BB_MAP_VZ, BB_MAP_SXX, BB_MAP_SYY, BB_MAP_SZZ,
    BB_MAP_SXY, BB_MAP_SXZ, BB_MAP_SYZ, BB_MAP_PRESSURE, BB_MAP_COUNT
};
#define BB_NCOEF 8 /* per-material row: M, G, L, B, tauL, tauS, 1/tau_sigma, K  (all / h) */

typedef struct bb_fdtd bb_fdtd;

typedef struct bb_fdtd_desc {
    int32_t n1, n2, n3;          /* global grid (MaterialMap.shape) */
    int32_t i0, i1;              /* planes owned by this handle */
    int32_t pml;                 /* NDelta (BabelIntegrationBASE.py:2350) */
    int32_t nmat;                /* rows of MaterialList */
    int32_t nsrc;                /* rows of SourceFunctions */
    int32_t nt_src;              /* columns of SourceFunctions (LengthSource) */
    int32_t steps;               /* time steps of the run */
    int32_t type_source;         /* TypeSource: 0/1 particle soft/hard, 2/3 stress soft/hard */
    int32_t sel_rms_peak;        /* SelRMSorPeak: 1 RMS, 2 peak, 3 both */
    uint32_t sel_maps_rms;       /* SelMapsRMSPeakList as a bit mask of bb_map */
    uint32_t sel_maps_sensor;    /* SelMapsSensorsList as a bit mask of bb_map */
    int32_t sensor_subsampling;  /* SensorSubSampling */
    int32_t sensor_start;        /* SensorStart */
    int32_t device;              /* CUDA ordinal */
    int32_t rank, nranks;        /* slab rank (0,1 on one GPU) */
    int32_t kernel_variant;      /* 0 = default (fastest); other values select debug variants */
    float mpml_ratio;            /* multi-axial PML: fraction of each axis' damping applied to the split parts of the two
                                    other axes (Meza-Fajardo & Papageorgiou 2008); 0 = classical split-field layer, which
                                    is unstable where a fluid-solid interface enters the layer; the host passes 0.1 */
    double dt;                   /* DT */
} bb_fdtd_desc;

typedef struct bb_fdtd_stats {
    double run_ms;               /* device time of the last bb_fdtd_run (CUDA events) */
    double stress_ms;            /* sum over launches of the stress kernel (profile mode) */
    double particle_ms;          /* sum over launches of the particle kernel (profile mode) */
    double pml_ms;               /* sum over launches of the PML kernels (profile mode) */
    double other_ms;             /* sources + sensors + halo (profile mode) */
    int64_t stress_launches, particle_launches, pml_launches, other_launches;
    int64_t steps_done;
    int64_t cells_local;         /* (i1-i0)*N2*N3 */
    int64_t device_bytes;        /* device memory held by the handle */
    int64_t nsamples;            /* sensor samples per sensor */
} bb_fdtd_stats;

/* ---- library / device ---- */
const char *bb_last_error(void);
const char *bb_version(void);
int bb_device_count(void);                                   /* <0 on error */
int bb_device_name(int device, char *out, int out_len);

/* page-locked host buffers for results (device-to-host copies into them run at link speed and skip
 * the first-touch page faults of fresh pageable memory); freed with bb_host_free */
int bb_host_alloc(int64_t bytes, void **out);
int bb_host_free(void *ptr);

/* host-side gather of a slab-decomposed run: out[rows[r]] = data[r] for nrows rows of row_bytes bytes each (the rows of
 * one slab's sensor traces placed into the whole-grid table in IndexSensorMap order).  Plain memcpy loop; exists so
 * that the per-GPU host threads of the Python layer gather in parallel without holding the interpreter lock. */
int bb_host_scatter_rows(void *out, const int64_t *rows, const void *data, int64_t nrows, int64_t row_bytes);
/* the same by runs of consecutive rows: rows [src_row[r], src_row[r] + nrows[r]) of data go to rows
 * [dst_row[r], dst_row[r] + nrows[r]) of out (one run per (j,k) line of a slab: slab.merge_sensor_runs) */
int bb_host_scatter_runs(void *out, const void *data, const int64_t *dst_row, const int64_t *src_row, const int64_t *nrows,
                         int64_t nruns, int64_t row_bytes);

/* flat index and value of every nonzero entry of a uint32 volume, in order, on a few host threads (the caller's SourceMap;
 * replaces np.flatnonzero in the host side of BabelIntegrationBASE.py:2338's call).  *count = number of nonzero entries;
 * nothing is written when it exceeds capacity. */
int bb_host_nonzero_u32(const uint32_t *a, int64_t n, int64_t *index, uint32_t *value, int64_t capacity, int64_t *count);

/* LZ4 block decoder (host) for the Blosc-LZ4 chunks of the HDF5 files the genuine H5pySimple writes (read here by
 * babelbrain_b200/h5mini.py when h5py is absent).  Returns the bytes written to dst, -1 for a corrupt block. */
long long bb_host_lz4_decompress(const unsigned char *src, long long n, unsigned char *dst, long long cap);

/* device memory of destroyed handles is kept per device for the next simulation of the same grid (a worker runs the forward,
 * back-propagation and refocus simulations in a row, BabelIntegrationBASE.py:2338-2428; at most BB_DEVICE_POOL_GB gigabytes,
 * default 32); this returns it to the driver.  device < 0: every device. */
int bb_release_cached_memory(int device);

/* ---- FDTD handle ---- */
int bb_fdtd_create(const bb_fdtd_desc *desc, bb_fdtd **out);
void bb_fdtd_destroy(bb_fdtd *h);
/* optional: run on the caller's CUDA stream (cudaStream_t as void*); NULL = library stream */
int bb_fdtd_set_stream(bb_fdtd *h, void *cuda_stream);
/* nmat x BB_NCOEF float table and the 2 x (pml+1) PML table: damping d at integer depth 0..pml and at half depth
 * xi + 0.5 (quadratic profile, d0 = ln(1/ReflectionLimit) 3 Vmax / (2 pml h)) */
int bb_fdtd_set_materials(bb_fdtd *h, const float *table, const float *pml_table);
/* uint32 label planes [max(i0-2,0), min(i1+2,n1)) ; reflector may be NULL (ReflectorMask=None) */
int bb_fdtd_set_maps(bb_fdtd *h, const uint32_t *material, const uint32_t *reflector);
/* source cells owned by this slab: global C-order linear cell index, 0-based source row, and the
 * Ox/Oy/Oz weights at those cells (BabelIntegrationBASE.py:2328-2335) */
int bb_fdtd_set_source_cells(bb_fdtd *h, int64_t ncells, const int64_t *cell, const int32_t *row,
                             const float *ox, const float *oy, const float *oz);
/* SourceFunctions as the caller holds it: (nsrc, nt_src), float64 or float32, row stride in
 * elements (BabelIntegrationSingle.py:335).  Converted and transposed on the device. */
int bb_fdtd_set_source_functions(bb_fdtd *h, const void *data, int is_f64, int64_t row_stride);
/* the same table, uploaded in chunks of time samples while bb_fdtd_run advances (each chunk is on the device before the
 * step that reads it; the first two before the first step).  `data` must stay valid and unchanged until bb_fdtd_run has
 * passed time step nt_src - 1 -- the caller of StaggeredFDTD_3D_with_relaxation holds it for the whole call anyway
 * (BabelIntegrationBASE.py:2338-2365).  Tables with so many rows that a chunk would hold < 64 samples are uploaded at once. */
int bb_fdtd_set_source_functions_streamed(bb_fdtd *h, const void *data, int is_f64, int64_t row_stride);
/* Alternative to bb_fdtd_set_source_functions for the continuous-wave sources every transducer model builds
 * (CreateSources, BabelIntegrationSingle.py:313-346: row s = |u0_s| sin(2 pi f t + angle(u0_s)), the first ramp samples
 * scaled by a raised cosine): the rows are evaluated in the source kernel from a_cos[s] = |u0_s| cos(angle), a_sin[s] =
 * |u0_s| sin(angle) (nsrc floats each) and the two time envelopes env_sin[n] = ramp(n) sin(2 pi f t_n), env_cos[n] =
 * ramp(n) cos(2 pi f t_n) (nt_src floats each, rounded from double by the caller).  No (nsrc, nt_src) table exists on
 * either side of the PCIe link (at 1 MHz / 1080^3 that table is 100 GB of float64 on the host). */
int bb_fdtd_set_source_tones(bb_fdtd *h, const float *a_cos, const float *a_sin, const float *env_sin, const float *env_cos);
/* sensors owned by this slab: global C-order linear cell index, in IndexSensorMap order */
int bb_fdtd_set_sensors(bb_fdtd *h, int64_t nsensors, const int64_t *cell);
/* Alternative to bb_fdtd_set_sensors: build the sensor table on the device from the caller's
 * SensorMap (BabelIntegrationBASE.py:2283-2290).  sensor_map = planes [i0,i1) of the (N1,N2,N3)
 * C-order uint32 volume; every non-zero voxel becomes a sensor, ordered like IndexSensorMap
 * (ascending 1-based Fortran-order linear index i + j*N1 + k*N1*N2 + 1, :2503-2511). */
int bb_fdtd_set_sensor_map(bb_fdtd *h, const uint32_t *sensor_map, int64_t *nsensors);
/* the 1-based Fortran-order indices of this slab's sensors, in table order; elem_bytes = 4 or 8 */
int bb_fdtd_get_sensor_index(bb_fdtd *h, void *out, int elem_bytes);
/* slab neighbours exchange halos with NCCL send/recv; id = 128-byte ncclUniqueId from rank 0 */
int bb_nccl_unique_id(char *out128);
int bb_fdtd_comm_init(bb_fdtd *h, const char *id128);        /* id128 == NULL: reuse this process's communicator of the same (device, rank, nranks) */
/* NVLink halo push (preferred over the NCCL exchange when every neighbour is reachable as CUDA peer memory):
 * each rank exports a descriptor of its slab, the host side hands every rank its neighbours' descriptors
 * (same process: raw peer pointers; another process: CUDA IPC handles), and from then on the boundary CTAs of each
 * half-step kernel store the two planes a neighbour needs straight into its halo planes and publish a sequence
 * number there; no separate exchange step, no NCCL call in the time loop. */
typedef struct bb_peer_info {
    int64_t pid;                 /* exporting process */
    int32_t device, nown;        /* CUDA ordinal, planes owned */
    int32_t n2, pitch;           /* transverse geometry (must match) */
    uint64_t v_ptr, s_ptr, flag_ptr;          /* device addresses in the exporting process */
    unsigned char v_ipc[64], s_ipc[64], flag_ipc[64];   /* cudaIpcMemHandle_t of the three allocations */
} bb_peer_info;
int bb_fdtd_peer_export(bb_fdtd *h, bb_peer_info *out);
/* lower / upper = descriptor of the rank owning the planes below / above this slab; NULL where there is none */
int bb_fdtd_peer_attach(bb_fdtd *h, const bb_peer_info *lower, const bb_peer_info *upper);
/* advance nsteps (<0: all remaining).  profile != 0 brackets each kernel with CUDA events. */
int bb_fdtd_run(bb_fdtd *h, int64_t nsteps, int profile);
int bb_fdtd_reset(bb_fdtd *h);                               /* zero state, step counter = 0 */
/* results: which = 0 RMS, 1 peak, 2 last field; out = (i1-i0, N2, N3) float32 */
int bb_fdtd_get_map(bb_fdtd *h, int which, int map_id, float *out);
/* out = (nsensors, nsamples) float32 for one selected sensor map */
int bb_fdtd_get_sensors(bb_fdtd *h, int map_id, float *out);
/* one slab of a multi-GPU run: the slab's rows go straight into the whole-grid (table_rows, nsamples) table `table` in
 * page-locked host memory, run by run (run u = the slab's rows [src_row[u], src_row[u] + nrows[u]) -> table rows from
 * dst_row[u]; the runs tile the slab's rows in order: slab.merge_sensor_runs).  *done = 0 and nothing written when `table`
 * is not page-locked: use bb_fdtd_get_sensors + bb_host_scatter_runs then.  Replaces the per-sensor placement the caller
 * would otherwise do on the host after gathering slabs (the reference has one GPU and no such step). */
int bb_fdtd_get_sensors_runs(bb_fdtd *h, int map_id, float *table, int64_t table_rows, const int64_t *dst_row,
                             const int64_t *src_row, const int64_t *nrows, int64_t nruns, int *done);
/* Phase / amplitude extraction of the sampled traces on the device -- what the caller's CalculatePhaseData does on the
 * host with an FFT over Sensor['Pressure'] (BabelIntegrationBASE.py:2498-2518): for every sensor of this slab, bin `bin`
 * of the DFT of its first nsamples_used samples (sum_n x[n] exp(-2 pi j bin n / nsamples_used)) times `scale`
 * (the caller's 2/nsamples, :2518), the angle of that bin and the largest sample, scattered to the sensor's voxel of
 * dense (i1-i0, N2, N3) volumes; voxels without a sensor are 0 (:2474-2476).  fourier_reim holds (re, im) pairs
 * (a numpy complex64 volume); phase and peak may be NULL.  Replaces the download of (nsensors, nsamples) traces plus
 * IndexSensorMap and the host FFT by two or three volume downloads. */
int bb_fdtd_get_phase_data(bb_fdtd *h, int map_id, int bin, int nsamples_used, float scale,
                           float *fourier_reim, float *phase, float *peak);
int bb_fdtd_get_stats(bb_fdtd *h, bb_fdtd_stats *out);
/* profiling aid (handle created with BB_CTA_TIMING=1 in the environment): per CTA of the most recent half-step launch
 * {start ns, end ns, blockIdx packed z<<40|y<<20|x, planes}; out holds 4*n uint64 */
int bb_fdtd_debug_cta_times(bb_fdtd *h, unsigned long long *out, int64_t n);

/* ---- Rayleigh integral ---- */
/* out[p] = j k /(2 pi) * sum_s ds[s] exp(Im(k) R)/R u0[s] exp(-j Re(k) R); host pointers;
 * center (nsrc,3), u0_reim (nsrc,2), rf (npts,3), out_reim (npts,2); max_distance<=0: no skip.
 * u0_step != 0: per-point source amplitudes, u0_reim is (npts*nsrc, 2). */
int bb_rayleigh_forward(float k_re, float k_im, int64_t nsrc, const float *center, const float *ds,
                        const float *u0_reim, int64_t npts, const float *rf, float *out_reim,
                        float max_distance, int64_t u0_step, int device, double *kernel_ms);

/* ---- bio-heat transfer (thermal step) ---- */
/* Pennes equation + CEM43 dose, explicit 7-point stencil; replaces BHTE / BHTEMultiplePressureFields
 * (ThermalModeling/CalculateTemperatureEffects.py:14, :365-395, :406, :439, :960-990).  Host pointers.
 * q: nfields x (n1,n2,n3) heat added per step where that field's beam is on (already scaled by dt and the duty cycle);
 * labels: (n1,n2,n3) uint32 material map; bh / perf: nmat conduction and perfusion coefficients per step;
 * temp / dose: (n1,n2,n3) in = initial state, out = state after total_steps; field_at_step[n] = index of the pressure field
 * heating during step n, -1 = beam off; monitor_slice (n1, n3, total_steps / nfactor_monitoring) receives the plane
 * j = sel_j every nfactor_monitoring-th step (NULL: none); monitor_points: (n1,n2,n3) map of 1-based point ids (NULL: none),
 * temp_points (npoints, total_steps) their temperature histories. */
int bb_bhte_run(int n1, int n2, int n3, int nmat, int nfields, const float *q, const uint32_t *labels, const float *bh,
                const float *perf, float *temp, float *dose, const int16_t *field_at_step, int64_t total_steps, float dt,
                float core_temp, int sel_j, int nfactor_monitoring, float *monitor_slice, const uint32_t *monitor_points,
                int64_t npoints, float *temp_points, int device, double *kernel_ms);

#ifdef __cplusplus
}
#endif
#endif
