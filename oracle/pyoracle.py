"""ctypes binding of the CPU oracle (oracle/gnss_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() (as the checker) and
bench.py's cpu_baseline / --impl reference legs.  The product (gnss-sdr-rs_b200/) never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c32 = np.complex64


class AcqResult(C.Structure):
    _fields_ = [("prn", C.c_uint8), ("code_phase_samples", C.c_uint64), ("code_phase_chips", C.c_float),
                ("carrier_freq", C.c_float), ("fs", C.c_float), ("mag_relative", C.c_float),
                ("sample_global_index", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class LoopFilter(C.Structure):
    _fields_ = [("tau1", C.c_float), ("tau2", C.c_float)]


class TrkChannel(C.Structure):
    _fields_ = [("id", C.c_uint8), ("prn", C.c_uint8), ("state", C.c_int32), ("code_row", C.c_int32),
                ("lost_counter", C.c_uint32), ("fs", C.c_float), ("next_sample_index", C.c_uint64),
                ("num_samples_per_code", C.c_uint64),
                ("carrier_freq", C.c_float), ("carrier_phase", C.c_float), ("carrier_error", C.c_float),
                ("carrier_nco", C.c_float), ("code_phase", C.c_float), ("code_error", C.c_float),
                ("code_nco", C.c_float), ("code_rate", C.c_float), ("i_prompt", C.c_float),
                ("q_prompt", C.c_float), ("pll_filter", LoopFilter), ("dll_filter", LoopFilter)]


class Ring(C.Structure):
    _fields_ = [("buffer", C.c_void_p), ("buf_size", C.c_size_t), ("mask", C.c_size_t), ("head", C.c_size_t)]


class Frontend(C.Structure):
    _fields_ = [("lut_re", C.c_float * 2048), ("lut_im", C.c_float * 2048), ("phase_accumulator", C.c_float),
                ("phase_step", C.c_float), ("bias_re", C.c_float * 8), ("bias_im", C.c_float * 8),
                ("alpha", C.c_float), ("con", C.c_float)]


class NavSync(C.Structure):
    _fields_ = [("flag_bit_sync", C.c_int32), ("frame_sync_ind", C.c_int32), ("sync_epoch", C.c_int32),
                ("n_bits", C.c_int32), ("bit_sync_buff", C.c_uint32 * 20), ("preamble_bit", C.c_int32),
                ("polarity", C.c_int32), ("ref_frame_sync", C.c_int32), ("ref_polarity", C.c_int32)]


class FineResult(C.Structure):
    _fields_ = [("fft_size", C.c_uint32), ("idx", C.c_uint32), ("mag", C.c_float), ("carrier_freq", C.c_float),
                ("ref_defined", C.c_int32)]


class AcqManager(C.Structure):
    _fields_ = [("mode", C.c_int)]


CELL_DTYPE = np.dtype([("peak", np.float32), ("argmax", np.uint32), ("sum8", np.float32)])


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    build()
    L = C.CDLL(os.path.join(_HERE, "libgnss_oracle.so"))
    vp, i32, f32, u64, sz = C.c_void_p, C.c_int, C.c_float, C.c_uint64, C.c_size_t
    sig = {
        "go_ca_code_chips": (i32, [i32, vp]),
        "go_ca_table": (vp, []),
        "go_generate_ca_code_samples": (i32, [i32, f32, f32, vp, i32]),
        "go_num_samples_per_code": (i32, [f32, f32]),
        "go_fft_plan_new": (vp, [i32, i32]), "go_fft_plan_free": (None, [vp]), "go_fft_process": (None, [vp, vp]),
        "go_fft64_plan_new": (vp, [i32, i32]), "go_fft64_plan_free": (None, [vp]),
        "go_fft64_process": (None, [vp, vp]),
        "go_fft_forward": (None, [i32, vp]), "go_fft_power_spectrum": (None, [i32, vp, vp]),
        "go_rfft_forward": (None, [i32, vp, vp]),
        "go_doppler_table": (f32, [f32, f32, f32, i32, vp]),
        "go_apply_doppler_shift": (None, [vp, vp, vp, i32]),
        "go_acq_worker_new": (vp, [i32, i32, f32]), "go_acq_worker_new_code": (vp, [i32, i32, f32, vp]),
        "go_acq_worker_free": (None, [vp]), "go_acq_worker_code_fft": (vp, [vp]),
        "go_acq_search": (i32, [vp, vp, vp, vp, i32, u64, i32, vp]),
        "go_acq_cells": (None, [vp, vp, vp, i32, i32, i32, vp, i32, vp]),
        "go_acq_bin_power": (None, [vp, vp, vp, i32, i32, vp, i32, vp]),
        "go_coh_rotators": (None, [f32, f32, i32, i32, vp]),
        "go_is_good_cell": (i32, [f32, f32, i32, f32]),
        "go_acq_decide": (i32, [vp, vp, i32, i32, i32, f32, u64, f32, vp, vp]),
        "go_two_peak_ratio": (f32, [vp, i32, i32, vp, vp]),
        "go_acq_search_all": (None, [vp, i32, vp, vp, vp, i32, u64, i32, i32, i32, vp, vp]),
        "go_acq_cells_all": (None, [vp, i32, vp, vp, i32, i32, i32, vp, i32, i32, vp]),
        "go_acq_manager_update_mode": (None, [vp, sz]),
        "go_acq_manager_pacing": (None, [vp, C.c_uint32, vp, vp]),
        "go_loop_filter_new": (LoopFilter, [f32, f32, f32]),
        "go_loop_filter_update": (f32, [vp, f32, f32, f32]),
        "go_trk_channel_init": (None, [vp, C.c_uint8, f32]), "go_trk_channel_start": (None, [vp, vp]),
        "go_trk_channel_reset": (None, [vp]), "go_trk_get_ca_chip": (f32, [vp, f32]),
        "go_trk_early_late": (None, [vp, vp, vp]), "go_trk_run_loop_filters": (None, [vp, vp]),
        "go_trk_do_work": (i32, [vp, vp, vp, vp]),
        "go_ring_init": (i32, [vp, sz]), "go_ring_free": (None, [vp]), "go_ring_write": (None, [vp, vp, sz]),
        "go_ring_head": (sz, [vp]), "go_ring_copy_to_slice": (None, [vp, sz, vp, sz]),
        "go_trk_update": (i32, [vp, vp, vp, vp, vp, vp]),
        "go_trk_run_all": (None, [vp, i32, vp, sz, i32, i32, vp]),
        "go_nav_bit_sync": (None, [vp, i32, i32, vp, vp, i32]),
        "go_frontend_init": (None, [vp, f32, f32]),
        "go_frontend_process_block": (None, [vp, vp, sz]),
        "go_fine_doppler": (i32, [vp, sz, vp, sz, f32, i32, i32, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------- convenience wrappers
def ca_table():
    L = lib()
    ptr = L.go_ca_table()
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int8)), shape=(32, 1023)).copy()


def ca_code_samples(prn, code_rate, fs):
    L = lib()
    n = L.go_num_samples_per_code(code_rate, fs)
    out = np.zeros(n, np.int8)
    L.go_generate_ca_code_samples(prn, code_rate, fs, _p(out), n)
    return out


def fft(x, inverse=False):
    L = lib()
    x = np.ascontiguousarray(x, dtype=c32).copy()
    p = L.go_fft_plan_new(len(x), int(inverse))
    L.go_fft_process(p, _p(x))
    L.go_fft_plan_free(p)
    return x


def fft64(x, inverse=False):
    L = lib()
    x = np.ascontiguousarray(x, dtype=np.complex128).copy()
    p = L.go_fft64_plan_new(len(x), int(inverse))
    L.go_fft64_process(p, _p(x))
    L.go_fft64_plan_free(p)
    return x


def doppler_table(f_if, f_d, fs, n):
    L = lib()
    t = np.zeros(n, c32)
    carr = L.go_doppler_table(f_if, f_d, fs, n, _p(t))
    return carr, t


def doppler_tables(f_if, dopplers, fs, n):
    tabs = np.zeros((len(dopplers), n), c32)
    carr = np.zeros(len(dopplers), np.float32)
    L = lib()
    for d, fd in enumerate(dopplers):
        carr[d] = L.go_doppler_table(f_if, float(fd), fs, n, _p(tabs[d]))
    return carr, tabs


def apply_doppler_shift(samples, table, out=None):
    L = lib()
    samples = np.ascontiguousarray(samples, c32)
    if out is None:
        out = np.zeros(len(samples), c32)
    L.go_apply_doppler_shift(_p(samples), _p(table), _p(out), len(samples))
    return out


def coh_rotators(carr, fs, n, n_coh):
    L = lib()
    carr = np.atleast_1d(np.asarray(carr, np.float32))
    rot = np.zeros((len(carr), n_coh), c32)
    for d in range(len(carr)):
        L.go_coh_rotators(float(carr[d]), fs, n, n_coh, _p(rot[d]))
    return rot


class AcqWorker:
    """AcquisitionWorker (do_acquisition.rs:118-239)."""

    def __init__(self, prn, fft_size, fs, code_samples=None):
        self.L = lib()
        self.prn, self.n, self.fs = prn, fft_size, fs
        if code_samples is None:
            self.h = self.L.go_acq_worker_new(prn, fft_size, fs)
        else:
            cs = np.ascontiguousarray(code_samples, np.int8)
            assert len(cs) == fft_size
            self.h = self.L.go_acq_worker_new_code(prn, fft_size, fs, _p(cs))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.go_acq_worker_free(self.h)
            self.h = None

    def code_fft(self):
        ptr = self.L.go_acq_worker_code_fft(self.h)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(self.n, 2)).copy().view(c32).ravel()

    def search_satellite(self, samples, tables, carr, local_tail, num_integrations):
        samples = np.ascontiguousarray(samples, c32)
        tables = np.ascontiguousarray(tables, c32)
        carr = np.ascontiguousarray(carr, np.float32)
        res = AcqResult()
        ok = self.L.go_acq_search(self.h, _p(samples), _p(tables), _p(carr), len(carr), local_tail,
                                  num_integrations, C.byref(res))
        return res.as_dict() if ok else None

    def cells(self, samples, tables, num_integrations, n_coh=1, rot=None, presum=0):
        samples = np.ascontiguousarray(samples, c32)
        tables = np.ascontiguousarray(tables, c32)
        D = tables.shape[0]
        out = np.zeros(D, CELL_DTYPE)
        rp = _p(np.ascontiguousarray(rot, c32)) if rot is not None else None
        self.L.go_acq_cells(self.h, _p(samples), _p(tables), D, num_integrations, n_coh, rp, presum, _p(out))
        return out

    def bin_power(self, samples, table, num_integrations, n_coh=1, rot=None, presum=0):
        samples = np.ascontiguousarray(samples, c32)
        out = np.zeros(self.n, np.float32)
        rp = _p(np.ascontiguousarray(rot, c32)) if rot is not None else None
        self.L.go_acq_bin_power(self.h, _p(samples), _p(np.ascontiguousarray(table, c32)), num_integrations, n_coh,
                                rp, presum, _p(out))
        return out


def acq_decide(cells, carr, prn, fft_size, fs, local_tail=0, threshold=7.0):
    L = lib()
    cells = np.ascontiguousarray(cells, CELL_DTYPE)
    carr = np.ascontiguousarray(carr, np.float32)
    res = AcqResult()
    b = C.c_int(-1)
    ok = L.go_acq_decide(_p(cells), _p(carr), len(carr), prn, fft_size, fs, local_tail, threshold, C.byref(res),
                         C.byref(b))
    if not ok:
        return None
    d = res.as_dict()
    d["bin"] = b.value
    return d


def acq_cells_all(workers, samples, tables, num_integrations, n_coh=1, rot=None, presum=0, n_threads=None):
    L = lib()
    samples = np.ascontiguousarray(samples, c32)
    tables = np.ascontiguousarray(tables, c32)
    D = tables.shape[0]
    arr = (C.c_void_p * len(workers))(*[w.h for w in workers])
    out = np.zeros((len(workers), D), CELL_DTYPE)
    rp = _p(np.ascontiguousarray(rot, c32)) if rot is not None else None
    L.go_acq_cells_all(arr, len(workers), _p(samples), _p(tables), D, num_integrations, n_coh, rp, presum,
                       n_threads or os.cpu_count(), _p(out))
    return out


def acq_search_all(workers, samples, tables, carr, local_tail, num_integrations, early_exit=True, n_threads=None):
    L = lib()
    samples = np.ascontiguousarray(samples, c32)
    tables = np.ascontiguousarray(tables, c32)
    carr = np.ascontiguousarray(carr, np.float32)
    arr = (C.c_void_p * len(workers))(*[w.h for w in workers])
    found = (C.c_int * len(workers))()
    results = (AcqResult * len(workers))()
    L.go_acq_search_all(arr, len(workers), _p(samples), _p(tables), _p(carr), len(carr), local_tail,
                        num_integrations, int(early_exit), n_threads or os.cpu_count(), found, results)
    return [results[i].as_dict() if found[i] else None for i in range(len(workers))]


def trk_channel(id_, fs):
    ch = TrkChannel()
    lib().go_trk_channel_init(C.byref(ch), id_, fs)
    return ch


def trk_start(ch, prn, carrier_freq, code_phase_chips, sample_global_index, fs, code_phase_samples=0, mag=0.0):
    r = AcqResult(prn, code_phase_samples, code_phase_chips, carrier_freq, fs, mag, sample_global_index)
    lib().go_trk_channel_start(C.byref(ch), C.byref(r))


def trk_early_late(ch, data):
    data = np.ascontiguousarray(data, c32).copy()
    out = np.zeros(6, np.float32)
    lib().go_trk_early_late(C.byref(ch), _p(data), _p(out))
    return out


def trk_do_work(ch, data):
    data = np.ascontiguousarray(data, c32).copy()
    out = np.zeros(6, np.float32)
    mp = C.c_uint8(0)
    msg = lib().go_trk_do_work(C.byref(ch), _p(data), _p(out), C.byref(mp))
    return out, msg, mp.value


def trk_run_all(channels, stream, n_epochs, n_threads=None, want_hist=True):
    """channels: ctypes array (TrkChannel * C). Returns prompt history [n_epochs, C, 2]."""
    stream = np.ascontiguousarray(stream, c32)
    n = len(channels)
    hist = np.zeros((n_epochs, n, 2), np.float32) if want_hist else None
    lib().go_trk_run_all(channels, n, _p(stream), len(stream), n_epochs, n_threads or os.cpu_count(),
                         _p(hist) if want_hist else None)
    return hist


def frontend(f_if, fs_in):
    f = Frontend()
    lib().go_frontend_init(C.byref(f), f_if, fs_in)
    return f


def frontend_process(f, samples):
    x = np.ascontiguousarray(samples, c32).copy()
    lib().go_frontend_process_block(C.byref(f), _p(x), len(x))
    return x


def nav_bit_sync(prompt_i, max_bits=4096):
    """prompt_i: [n_epochs] float32 of ONE channel -> (NavSync, bits)."""
    x = np.ascontiguousarray(prompt_i, np.float32)
    st = NavSync()
    bits = np.zeros(max_bits, np.int8)
    lib().go_nav_bit_sync(_p(x), len(x), 1, C.byref(st), _p(bits), max_bits)
    return st, bits[:min(st.n_bits, max_bits)].copy()


def fine_doppler(long_samples, code1023, code_phase, fs, long_ms=11, is_complex=True, want_mag=False):
    """finer_doppler (acquisition_bk.rs:215-302) -> (FineResult, mag or None); None if the recording is too short."""
    x = np.ascontiguousarray(long_samples, c32)
    code = np.ascontiguousarray(code1023, np.int8)
    res = FineResult()
    mag = None
    if want_mag:
        n_code = int(round(fs / 1000.0))
        p2 = 1
        while p2 < (long_ms - 1) * n_code:
            p2 <<= 1
        mag = np.zeros(8 * p2, np.float32)
    rc = lib().go_fine_doppler(_p(x), len(x), _p(code), int(code_phase), fs, int(long_ms), int(bool(is_complex)),
                               C.byref(res), _p(mag) if want_mag else None)
    if rc != 0:
        return None, None
    return res, mag
