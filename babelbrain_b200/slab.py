"""
Slab decomposition of the FDTD domain along axis 0 (the slowest-varying axis of the caller's
C-order volumes, so every halo is a set of contiguous planes).  Pure host logic: used by the CUDA
path (one process per GPU, halos exchanged inside libbabelb200.so with NCCL send/recv) and
exercised on CPU with world_size-2 gloo tests.  The reference has no distributed layer at all
(SURVEY.md section 2b); the single-GPU result is the parity oracle of the multi-GPU one.
"""
import numpy as np

HALO = 2  # planes: the 4th-order staggered stencil reaches 2 cells one way and 1 the other


class SlabPlan:
    def __init__(self, n1, nranks, pml=12, min_planes=4):
        n1, nranks = int(n1), int(nranks)
        if nranks < 1:
            raise ValueError('nranks must be >= 1')
        if nranks > 1 and n1 // nranks < min_planes:
            raise ValueError('too many ranks: %d planes over %d ranks (need >= %d each)' % (n1, nranks, min_planes))
        self.n1, self.nranks, self.pml = n1, nranks, int(pml)
        base, rem = divmod(n1, nranks)
        sizes = [base + (1 if r < rem else 0) for r in range(nranks)]
        self.bounds = np.concatenate([[0], np.cumsum(sizes)]).astype(int)

    def owned(self, rank):
        return int(self.bounds[rank]), int(self.bounds[rank + 1])

    def with_halo(self, rank):
        i0, i1 = self.owned(rank)
        return max(i0 - HALO, 0), min(i1 + HALO, self.n1)

    def neighbours(self, rank):
        return (rank - 1 if rank > 0 else None), (rank + 1 if rank < self.nranks - 1 else None)

    def rank_of_plane(self, i):
        return int(np.searchsorted(self.bounds, i, side='right') - 1)

    def halo_bytes_per_half_step(self, n2, pitch):
        """bytes one interior rank sends per half-step: 3 fields x 2 planes x 2 neighbours, fp32."""
        return 3 * HALO * 2 * n2 * pitch * 4


def halo_exchange_plan(rank, nranks, nown):
    """The exchange libbabelb200.so performs after each half-step (fdtd.cu: halo_exchange), as a list of
    (op, peer, first_local_plane, nplanes) over the local array of nown + 2*HALO planes: the first /
    last HALO owned planes go to the lower / upper neighbour, whose copies land in this rank's halo
    planes.  Three fields move per half-step: (Sxx, Sxy, Sxz) after the stress update -- the stresses
    the particle update differentiates along i -- and (Vx, Vy, Vz) after the particle update."""
    ops = []
    if rank > 0:
        ops.append(('send', rank - 1, HALO, HALO))
        ops.append(('recv', rank - 1, 0, HALO))
    if rank < nranks - 1:
        ops.append(('send', rank + 1, nown, HALO))
        ops.append(('recv', rank + 1, nown + HALO, HALO))
    return ops


def sensor_rows_of_slab(index_sensor_map, shape, i0, i1):
    """Rows of the global sensor table (IndexSensorMap order: 1-based Fortran linear index,
    BabelIntegrationBASE.py:2503-2511) whose voxel lies in planes [i0,i1)."""
    n1 = shape[0]
    i = (np.asarray(index_sensor_map).astype(np.int64) - 1) % n1
    return np.flatnonzero((i >= i0) & (i < i1))


def assemble_maps(shape, plan, slabs):
    """slabs: list over ranks of (i1-i0, N2, N3) arrays -> full (N1,N2,N3) volume."""
    out = np.empty(shape, np.float32)
    for r, a in enumerate(slabs):
        i0, i1 = plan.owned(r)
        out[i0:i1] = a
    return out


def assemble_sensors(nsensors, nsamples, rows_per_rank, data_per_rank):
    out = np.zeros((nsensors, nsamples), np.float32)
    for rows, data in zip(rows_per_rank, data_per_rank):
        out[rows] = data
    return out


def merge_sensor_tables(local_tables, n1, n23):
    """Global sensor table of a slab-decomposed run.  local_tables: per slab, the ascending 1-based
    Fortran-order indices (i + j*N1 + k*N1*N2 + 1, BabelIntegrationBASE.py:2503-2511) of the sensors in
    that slab's planes, as the device builds them.  Slabs cut axis 0, the fastest axis of that index, so
    within one (j,k) line the slabs' entries follow each other in rank order.  Returns (IndexSensorMap of
    the whole grid, list over slabs of the rows their sensors occupy in it) in O(sensors) without sorting."""
    lines, counts = [], []
    for t in local_tables:
        ln = (np.asarray(t).astype(np.int64) - 1) // int(n1)
        lines.append(ln)
        counts.append(np.bincount(ln, minlength=int(n23)).astype(np.int64))
    total = np.sum(counts, axis=0) if counts else np.zeros(int(n23), np.int64)
    before = np.cumsum(total) - total
    nall = int(total.sum())
    dtype = np.asarray(local_tables[0]).dtype if local_tables else np.uint32
    index = np.empty(nall, dtype)
    rows = []
    for t, ln, c in zip(local_tables, lines, counts):
        first = np.cumsum(c) - c
        r = before[ln] + (np.arange(ln.size, dtype=np.int64) - first[ln])
        index[r] = t
        rows.append(r)
        before = before + c
    return index, rows


def merge_sensor_runs(local_tables, n1, n23):
    """merge_sensor_tables without any per-sensor work: the entries of one (j,k) line are consecutive in a slab's
    table and consecutive in the global one, so the merge is a list of runs per slab -- (dst_row, src_row, nrows), one
    run per non-empty line -- found with one binary search per line.  Returns (total number of sensors, runs per slab);
    bb_host_scatter_runs (or expand_runs + fancy indexing) places table entries and trace rows with them."""
    n1, n23 = int(n1), int(n23)
    bounds, counts = [], []
    for t in local_tables:
        t = np.asarray(t)
        edges = np.arange(n23 + 1, dtype=np.int64) * n1 + 1                  # first 1-based index of every line
        if t.dtype.itemsize < 8 and n1 * n23 + 1 < 2 ** (8 * t.dtype.itemsize):
            edges = edges.astype(t.dtype)                                     # no up-cast copy of the table
        b = np.searchsorted(t, edges).astype(np.int64)
        bounds.append(b)
        counts.append(np.diff(b))
    total = np.sum(counts, axis=0) if counts else np.zeros(n23, np.int64)
    line_start = np.cumsum(total) - total
    runs, before = [], line_start
    for b, c in zip(bounds, counts):
        m = c > 0
        runs.append((np.ascontiguousarray(before[m]), np.ascontiguousarray(b[:-1][m]), np.ascontiguousarray(c[m])))
        before = before + c
    return int(total.sum()), runs


def expand_runs(run):
    """Global row of every entry of a slab's table, from its runs (the rows merge_sensor_tables returns)."""
    dst, src, n = run
    total = int(n.sum())
    return np.arange(total, dtype=np.int64) + np.repeat(dst - src, n)
