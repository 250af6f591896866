"""
Continuous-wave source tables without the table.

Every transducer model of the reference builds SourceFunctions the same way (CreateSources,
TranscranialModeling/BabelIntegrationSingle.py:313-346, repeated in the ANNULAR_ARRAY / CONCAVE_PHASEDARRAY / DOME files):
row s = |u0_s| sin(2 pi f t + angle(u0_s)) on t = arange(0, LengthSource + dt, dt), the first ramp samples scaled by a
raised cosine.  That is a dense float64 (Nsrc, Nt) matrix -- 0.95 GB for CTX-500, 100 GB for the 1 MHz 1080^3 case -- whose
information is two numbers per source.  CWSourceFunctions carries those numbers; passed as the SourceFunctions argument of
PropagationModel.StaggeredFDTD_3D_with_relaxation it makes the source kernel evaluate the rows in place
(bb_fdtd_set_source_tones), and .dense() returns the reference's matrix for any other consumer.
"""
import numpy as np


class CWSourceFunctions:
    ndim = 2
    dtype = np.dtype(np.float64)

    def __init__(self, amplitude, phase, Frequency, TemporalStep, TimeSimulation, ramp_length=4):
        self.amplitude = np.ascontiguousarray(amplitude, dtype=np.float64).reshape(-1)
        self.phase = np.ascontiguousarray(phase, dtype=np.float64).reshape(-1)
        if self.amplitude.shape != self.phase.shape:
            raise ValueError('amplitude and phase must have one entry per source')
        self.Frequency, self.TemporalStep = float(Frequency), float(TemporalStep)
        # BabelIntegrationSingle.py:315-316
        length = np.floor(TimeSimulation / (1.0 / self.Frequency)) * 1 / self.Frequency
        self.time = np.arange(0, length + self.TemporalStep, self.TemporalStep)
        # :319-324
        ramp_points = int(np.round(ramp_length / self.Frequency / self.TemporalStep))
        self.ramp = (-np.cos(np.arange(0, np.pi, np.pi / ramp_points)) + 1) * 0.5
        self.shape = (self.amplitude.size, self.time.size)

    @classmethod
    def from_source_plane(cls, SourceMapRayleigh, Frequency, TemporalStep, TimeSimulation, ramp_length=4):
        """Amplitudes and phases of the non-zero pixels of a complex source plane, in the row order CreateSources
        assigns (np.where order, :329, :338-345).  Returns (CWSourceFunctions, SourceMask uint32 with the 1-based rows)."""
        u = np.asarray(SourceMapRayleigh)
        ii, jj = np.where(np.abs(u) > 0)
        mask = np.zeros(u.shape, np.uint32)
        mask[ii, jj] = np.arange(1, ii.size + 1, dtype=np.uint32)
        u0 = u[ii, jj]
        return cls(np.abs(u0), np.angle(u0), Frequency, TemporalStep, TimeSimulation, ramp_length), mask

    def dense(self):
        """The reference's PulseSource matrix (:336-343), float64 (Nsrc, Nt)."""
        P = self.amplitude[:, None] * np.sin(2 * np.pi * self.Frequency * self.time[None, :] + self.phase[:, None])
        nr = min(len(self.ramp), P.shape[1])
        P[:, :nr] *= self.ramp[None, :nr]
        return P

    def __array__(self, dtype=None, copy=None):
        d = self.dense()
        return d if dtype is None else d.astype(dtype)

    def tone_tables(self):
        """What the device needs: per source (A cos phi, A sin phi), per time step (ramp sin wt, ramp cos wt), float32
        rounded from float64."""
        env = np.ones(self.time.size)
        nr = min(len(self.ramp), env.size)
        env[:nr] = self.ramp[:nr]
        wt = 2 * np.pi * self.Frequency * self.time
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)      # noqa: E731
        return (f32(self.amplitude * np.cos(self.phase)), f32(self.amplitude * np.sin(self.phase)),
                f32(env * np.sin(wt)), f32(env * np.cos(wt)))
