"""Host-side mirror of the reference crate's acquisition API over libgnss_b200.

Names and argument meaning follow src/acquisition/do_acquisition.rs and doppler_shift.rs:
DopplerShiftTable::new (doppler_shift.rs:11-21), AcquisitionWorker::{new, search_satellite}
(do_acquisition.rs:131-226), AcquisitionResult (:93-102), AcquisitionManager (:39-74).  The batched
engine replaces the rayon loop over 32 workers (:302-313) with one fused launch.
"""
import ctypes as C

import numpy as np

from . import _ffi

FREQ_SEARCH_ACQUISITION_HZ = 14e3  # do_acquisition.rs:20
FREQ_SEARCH_STEP_HZ = 500          # :21
PRN_SEARCH_ACQUISITION_TOTAL = 32  # :22
LONG_SAMPLES_LENGTH = 10           # :23
GPS_L1_CA_CODE_RATE_CHIPS_PER_S = np.float32(1.023e6)


def fft_size_for(fs):
    """round(fs / (code_rate / 1023)) as at do_acquisition.rs:249-251."""
    return _ffi.lib().gb_num_samples_per_code(float(GPS_L1_CA_CODE_RATE_CHIPS_PER_S), float(fs))


def reference_doppler_grid():
    """The 29 bins of run(): -7000 + 500*i (do_acquisition.rs:248-262)."""
    cap = int(np.uint16(FREQ_SEARCH_ACQUISITION_HZ) // FREQ_SEARCH_STEP_HZ) + 1
    return [np.float32(-FREQ_SEARCH_ACQUISITION_HZ / 2.0 + i * FREQ_SEARCH_STEP_HZ) for i in range(cap)]


class AcquisitionManager:
    """do_acquisition.rs:39-74 (host-only pacing logic)."""
    COLD, WARM, STEADY = 0, 1, 2

    def __init__(self):
        self.mode = self.COLD

    def update_mode(self, tracked_count):
        self.mode = self.COLD if tracked_count == 0 else (self.WARM if tracked_count <= 4 else self.STEADY)

    def get_pacing_and_list(self, active_prns):
        interval, size = {self.COLD: (500, 32), self.WARM: (1000, 8), self.STEADY: (2000, 5)}[self.mode]
        cands = [p for p in range(1, PRN_SEARCH_ACQUISITION_TOTAL + 1) if p not in active_prns][:size]
        mask = 0
        for p in cands:
            mask |= 1 << (p - 1)
        return interval, mask


class AcquisitionEngine:
    """All AcquisitionWorkers of one receiver on one GPU."""

    def __init__(self, handle, fft_size, fs, n_prn=32, codes=None):
        self.hd = handle
        self.n, self.fs, self.n_prn = int(fft_size), float(fs), int(n_prn)
        cp = None
        if codes is not None:
            codes = np.ascontiguousarray(codes, np.int8)
            assert codes.shape == (n_prn, fft_size)
            cp = _ffi.ptr(codes)
        handle.call("gb_acq_configure", self.n, self.fs, self.n_prn, cp)
        self.carr = None
        self.n_coh = 1

    # DopplerShiftTable::new for a list of Doppler offsets, built on the device
    def make_doppler_tables(self, f_if, dopplers):
        d = np.ascontiguousarray(dopplers, np.float32)
        carr = np.zeros(len(d), np.float32)
        self.hd.call("gb_acq_make_doppler_tables", float(f_if), _ffi.ptr(d), len(d), _ffi.ptr(carr))
        self.carr = carr
        return carr

    # caller-built tables (the reference's pub `table` / `doppler_freq_hz` fields)
    def set_doppler_tables(self, tables, carr):
        t = np.ascontiguousarray(tables, np.complex64)
        c = np.ascontiguousarray(carr, np.float32)
        assert t.shape == (len(c), self.n)
        self.hd.call("gb_acq_set_doppler_tables", _ffi.ptr(t), _ffi.ptr(c), len(c))
        self.carr = c.copy()

    def get_doppler_tables(self):
        t = np.zeros((len(self.carr), self.n), np.complex64)
        self.hd.call("gb_acq_get_doppler_tables", _ffi.ptr(t), None)
        return t

    def set_coherent(self, n_coh):
        self.hd.call("gb_acq_set_coherent", int(n_coh))
        self.n_coh = int(n_coh)

    def set_doppler_aliasing(self, on):
        """Share one forward spectrum between Doppler bins a whole number of FFT bins apart; off (the default) = every
        bin runs the reference's own wipe-off table (include/gnss_b200.h, gb_acq_set_doppler_aliasing)."""
        self.hd.call("gb_acq_set_doppler_aliasing", 1 if on else 0)

    def forward_bins(self):
        """Forward spectra per group the next shared-chain search computes."""
        return int(self.hd.L.gb_acq_forward_bins(self.hd.h))

    def set_mode(self, mode):
        """_ffi.GB_ACQ_FUSED (single kernel), _ffi.GB_ACQ_SHARED (forward path shared by all PRNs, default) or
        _ffi.GB_ACQ_SHARED_PLAIN (the shared chain with the generic inverse kernel at every size, A/B)."""
        self.hd.call("gb_acq_set_mode", int(mode))

    def set_detector(self, threshold=7.0, samples_per_chip=0):
        self.hd.call("gb_acq_set_detector", float(threshold), int(samples_per_chip))

    def _enable(self, enable):
        return None if enable is None else np.ascontiguousarray(enable, np.uint8)

    def _chunk(self, samples, num_integrations, n_samples):
        """Host chunk + its length in samples.  `samples` is an array, or the integer address of a (pinned) host buffer
        whose length the caller states in n_samples; the library refuses a buffer shorter than K * fft_size."""
        if isinstance(samples, int):
            if n_samples is None:
                raise ValueError("a raw host address needs n_samples")
            return samples, int(n_samples)
        x = np.ascontiguousarray(samples, np.complex64)
        return x, int(x.size)

    def search_cells(self, samples, num_integrations, prn_mask=0xFFFFFFFF, enable=None, n_samples=None):
        """Full PRN x Doppler grid -> structured array [n_prn, D] of {peak, argmax, sum8, peak2}."""
        x, n = self._chunk(samples, num_integrations, n_samples)
        cells = np.zeros((self.n_prn, len(self.carr)), _ffi.CELL_DTYPE)
        en = self._enable(enable)
        self.hd.call("gb_acq_search_cells", _ffi.ptr(x), n, int(num_integrations), int(prn_mask), _ffi.ptr(en),
                     _ffi.ptr(cells))
        return cells

    def search_cells_ring(self, local_tail, num_integrations, prn_mask=0xFFFFFFFF, enable=None, want_cells=True):
        cells = np.zeros((self.n_prn, len(self.carr)), _ffi.CELL_DTYPE) if want_cells else None
        en = self._enable(enable)
        self.hd.call("gb_acq_search_cells_ring", int(local_tail), int(num_integrations), int(prn_mask), _ffi.ptr(en),
                     _ffi.ptr(cells))
        return cells

    def search(self, samples, num_integrations, local_tail=0, prn_mask=0xFFFFFFFF, enable=None, n_samples=None):
        """search_satellite for every selected PRN: list of AcquisitionResult dicts or None."""
        x, n = self._chunk(samples, num_integrations, n_samples)
        res = (_ffi.AcqResult * self.n_prn)()
        en = self._enable(enable)
        self.hd.call("gb_acq_search", _ffi.ptr(x), n, int(num_integrations), int(local_tail), int(prn_mask),
                     _ffi.ptr(en), res)
        return [r.as_dict() if r.found else None for r in res]

    # asynchronous pair on one handle (two slots): the upload of one search overlaps the inverse kernel of the other
    def search_enqueue(self, samples, num_integrations, slot, local_tail=0, prn_mask=0xFFFFFFFF, enable=None, n_samples=None):
        x, n = self._chunk(samples, num_integrations, n_samples)
        self._keep = getattr(self, "_keep", {})
        self._keep[slot] = x   # the buffer must outlive the call
        en = self._enable(enable)
        self.hd.call("gb_acq_search_enqueue", _ffi.ptr(x), n, int(num_integrations), int(local_tail), int(prn_mask),
                     _ffi.ptr(en), int(slot))

    def search_wait(self, slot, want_cells=False, raw=None):
        res = raw if raw is not None else (_ffi.AcqResult * self.n_prn)()
        cells = np.zeros((self.n_prn, len(self.carr)), _ffi.CELL_DTYPE) if want_cells else None
        self.hd.call("gb_acq_search_wait", int(slot), res, _ffi.ptr(cells))
        getattr(self, "_keep", {}).pop(slot, None)
        if raw is not None:
            return raw
        out = [r.as_dict() if r.found else None for r in res]
        return (out, cells) if want_cells else out

    def search_ring(self, local_tail, num_integrations, prn_mask=0xFFFFFFFF, enable=None):
        res = (_ffi.AcqResult * self.n_prn)()
        en = self._enable(enable)
        self.hd.call("gb_acq_search_ring", int(local_tail), int(num_integrations), int(prn_mask), _ffi.ptr(en), res)
        return [r.as_dict() if r.found else None for r in res]

    def search_ring_raw(self, local_tail, num_integrations, prn_mask=0xFFFFFFFF, enable=None, out=None):
        """search_ring without the per-result Python objects: fills and returns a (gb_acq_result * n_prn) ctypes array
        (the form the multi-GPU gather ships as raw bytes)."""
        res = out if out is not None else (_ffi.AcqResult * self.n_prn)()
        en = self._enable(enable)
        self.hd.call("gb_acq_search_ring", int(local_tail), int(num_integrations), int(prn_mask), _ffi.ptr(en), res)
        return res

    def bin_power(self, samples, num_integrations, prn, doppler_bin):
        x = np.ascontiguousarray(samples, np.complex64)
        out = np.zeros(self.n, np.float32)
        self.hd.call("gb_acq_bin_power", _ffi.ptr(x), int(x.size), int(num_integrations), int(prn), int(doppler_bin),
                     _ffi.ptr(out))
        return out

    def last_kernel_ms(self):
        return float(self.hd.L.gb_acq_last_kernel_ms(self.hd.h))


def decide(cells_row, carr, prn, fft_size, fs, local_tail=0, threshold=7.0):
    """search_satellite's decision on one PRN's cells (gb_acq_decide)."""
    cells_row = np.ascontiguousarray(cells_row, _ffi.CELL_DTYPE)
    carr = np.ascontiguousarray(carr, np.float32)
    r = _ffi.AcqResult()
    _ffi.check(_ffi.lib().gb_acq_decide(_ffi.ptr(cells_row), _ffi.ptr(carr), len(carr), int(prn), int(fft_size),
                                        float(fs), int(local_tail), float(threshold), C.byref(r)), "gb_acq_decide")
    return r.as_dict() if r.found else None


LEGACY_LONG_SAMPLES_LENGTH = 11    # acquisition_bk.rs:21 (ms)


def finer_doppler(handle, long_samples, requests, fs, long_ms=LEGACY_LONG_SAMPLES_LENGTH, is_complex=True, codes1023=None,
                  want_mag=False):
    """finer_doppler (acquisition_bk.rs:215-302) for a batch of acquired satellites on one recording.

    long_samples: complex64 array (host), or an int = absolute index into the device ring (then `n_long` = long_ms*N
    samples are read from it).  requests: iterable of (prn, code_phase_samples).  Returns a FINE_RES_DTYPE array
    (fft_size, idx, mag, carrier_freq, ref_defined) and, if want_mag, the [n_req, fft_size] magnitudes.
    """
    req = np.zeros(len(requests), _ffi.FINE_REQ_DTYPE)
    for i, (prn, cp) in enumerate(requests):
        req[i]["prn"], req[i]["code_phase"] = int(prn), int(cp)
    out = np.zeros(len(req), _ffi.FINE_RES_DTYPE)
    codes = None if codes1023 is None else np.ascontiguousarray(codes1023, np.int8)
    if codes is not None:
        assert codes.shape == (len(req), 1023)
    if isinstance(long_samples, (int, np.integer)):
        n_long = int(long_ms) * fft_size_for(fs)
        handle.call("gb_acq_fine_doppler_ring", int(long_samples), n_long, float(fs), int(long_ms), int(bool(is_complex)),
                    _ffi.ptr(req), len(req), _ffi.ptr(codes), _ffi.ptr(out))
        return out
    x = np.ascontiguousarray(long_samples, np.complex64)
    mag = None
    if want_mag:
        use = (int(long_ms) - 1) * fft_size_for(fs)
        p2 = 1
        while p2 < use:
            p2 <<= 1
        mag = np.zeros((len(req), 8 * p2), np.float32)
    handle.call("gb_acq_fine_doppler", _ffi.ptr(x), len(x), float(fs), int(long_ms), int(bool(is_complex)), _ffi.ptr(req),
                len(req), _ffi.ptr(codes), _ffi.ptr(out), _ffi.ptr(mag))
    return (out, mag) if want_mag else out


class FFT:
    """fft.rs:5-30 FFT<T>, T = f32 (complex64 input) or f64 (complex128 input); any length."""

    def __init__(self, handle, length):
        self.hd, self.len = handle, int(length)

    def execute(self, x, inverse=False):
        f64 = np.asarray(x).dtype == np.complex128
        x = np.ascontiguousarray(x, np.complex128 if f64 else np.complex64)
        batch = x.size // self.len
        out = np.zeros_like(x)
        self.hd.call("gb_fft_c2c_f64" if f64 else "gb_fft_c2c", self.len, int(inverse), _ffi.ptr(x), _ffi.ptr(out), batch)
        return out

    def power_spectrum(self, x):
        f64 = np.asarray(x).dtype == np.complex128
        x = np.ascontiguousarray(x, np.complex128 if f64 else np.complex64)
        out = np.zeros(x.shape, np.float64 if f64 else np.float32)
        self.hd.call("gb_fft_power_spectrum_f64" if f64 else "gb_fft_power_spectrum", self.len, _ffi.ptr(x), _ffi.ptr(out),
                     x.size // self.len)
        return out


class RealFFT:
    """fft.rs:32-56 RealFFT<T>, T = f32 or f64 (by the input's dtype); any length."""

    def __init__(self, handle, length):
        self.hd, self.len = handle, int(length)

    def _run(self, x, power):
        f64 = np.asarray(x).dtype == np.float64
        x = np.ascontiguousarray(x, np.float64 if f64 else np.float32)
        batch = x.size // self.len
        if power:
            out = np.zeros((batch, self.len // 2 + 1), np.float64 if f64 else np.float32)
            name = "gb_rfft_power_spectrum_f64" if f64 else "gb_rfft_power_spectrum"
        else:
            out = np.zeros((batch, self.len // 2 + 1), np.complex128 if f64 else np.complex64)
            name = "gb_rfft_f64" if f64 else "gb_rfft"
        self.hd.call(name, self.len, _ffi.ptr(x), _ffi.ptr(out), batch)
        return out[0] if x.ndim == 1 else out

    def execute(self, x):
        return self._run(x, False)

    def power_spectrum(self, x):
        return self._run(x, True)
