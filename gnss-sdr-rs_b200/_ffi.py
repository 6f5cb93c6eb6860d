"""ctypes binding of libgnss_b200.so (include/gnss_b200.h).  Loud failure, no fallback."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgnss_b200.so")

GB_OK, GB_EINVAL, GB_ENODEVICE, GB_ECUDA, GB_EUNSUPPORTED, GB_ESTATE, GB_ENOMEM, GB_ERANGE, GB_ENCCL = 0, -1, -2, -3, -4, -5, -6, -7, -8
GB_TRK_IDLE, GB_TRK_TRACKING = 0, 1
GB_TRK_FAST, GB_TRK_ORDERED = 0, 1
GB_ACQ_FUSED, GB_ACQ_SHARED, GB_ACQ_SHARED_PLAIN = 0, 1, 2


class GnssB200Error(RuntimeError):
    def __init__(self, code, what, detail=""):
        self.code = code
        super().__init__("%s failed: %s (%d)%s" % (what, _strerror(code), code, (" -- " + detail) if detail else ""))


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("ring_capacity", C.c_uint64), ("flags", C.c_uint32)]


class AcqResult(C.Structure):
    _fields_ = [("prn", C.c_uint8), ("found", C.c_uint8), ("doppler_bin", C.c_int16),
                ("code_phase_samples", C.c_uint64), ("code_phase_chips", C.c_float), ("carrier_freq", C.c_float),
                ("fs", C.c_float), ("mag_relative", C.c_float), ("sample_global_index", C.c_uint64),
                ("metric", C.c_float), ("peak_ratio", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class TrkChannel(C.Structure):
    _fields_ = [("id", C.c_uint8), ("prn", C.c_uint8), ("state", C.c_uint8), ("code_row", C.c_uint8),
                ("lost_counter", C.c_uint32), ("fs", C.c_float), ("epochs_done", C.c_uint32),
                ("next_sample_index", C.c_uint64), ("num_samples_per_code", C.c_uint64),
                ("carrier_freq", C.c_float), ("carrier_phase", C.c_float), ("carrier_error", C.c_float),
                ("carrier_nco", C.c_float), ("code_phase", C.c_float), ("code_error", C.c_float),
                ("code_nco", C.c_float), ("code_rate", C.c_float), ("i_prompt", C.c_float), ("q_prompt", C.c_float),
                ("pll_tau1", C.c_float), ("pll_tau2", C.c_float), ("dll_tau1", C.c_float), ("dll_tau2", C.c_float)]


NAV_DTYPE = np.dtype([("flag_bit_sync", np.int32), ("frame_sync_ind", np.int32), ("sync_epoch", np.int32),
                      ("n_bits", np.int32), ("bit_sync_buff", np.uint32, (20,)), ("preamble_bit", np.int32),
                      ("polarity", np.int32), ("ref_frame_sync", np.int32), ("ref_polarity", np.int32)])
FINE_REQ_DTYPE = np.dtype([("prn", np.uint8), ("reserved", np.uint8, (3,)), ("code_phase", np.uint32)])
FINE_RES_DTYPE = np.dtype([("fft_size", np.uint32), ("idx", np.uint32), ("mag", np.float32), ("carrier_freq", np.float32),
                           ("ref_defined", np.int32)])
CELL_DTYPE = np.dtype([("peak", np.float32), ("argmax", np.uint32), ("sum8", np.float32), ("peak2", np.float32)])
CORR_DTYPE = np.dtype([("i_p", np.float32), ("q_p", np.float32), ("i_e", np.float32), ("q_e", np.float32),
                       ("i_l", np.float32), ("q_l", np.float32)])

_vp, _i32, _f32, _u64, _u32 = C.c_void_p, C.c_int, C.c_float, C.c_uint64, C.c_uint32

# every symbol include/gnss_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "gb_strerror": (C.c_char_p, [_i32]),
    "gb_last_cuda_error": (C.c_char_p, [_vp]),
    "gb_version": (_i32, []),
    "gb_device_count": (_i32, []),
    "gb_tuning_set": (_i32, [C.c_char_p, _i32]),
    "gb_tuning_get": (_i32, [C.c_char_p, _i32]),
    "gb_create": (_i32, [_vp, _vp]),
    "gb_destroy": (_i32, [_vp]),
    "gb_synchronize": (_i32, [_vp]),
    "gb_ca_code_chips": (_i32, [_i32, _vp]),
    "gb_num_samples_per_code": (_i32, [_f32, _f32]),
    "gb_generate_ca_code_samples": (_i32, [_i32, _f32, _f32, _vp, _i32]),
    "gb_ring_create": (_i32, [_vp, _u64]),
    "gb_ring_write": (_i32, [_vp, _vp, _u64]),
    "gb_ring_write_i8": (_i32, [_vp, _vp, _u64]),
    "gb_ring_head": (_u64, [_vp]),
    "gb_ring_copy_to_slice": (_i32, [_vp, _u64, _vp, _u64]),
    "gb_ring_reset": (_i32, [_vp]),
    "gb_frontend_configure": (_i32, [_vp, _f32, _f32]),
    "gb_frontend_write": (_i32, [_vp, _vp, _u64]),
    "gb_frontend_state": (_i32, [_vp, _vp]),
    "gb_frontend_set_mode": (_i32, [_vp, _i32]),
    "gb_frontend_orbit": (_i32, [_f32, _f32, _vp, _vp, _vp, _u64]),
    "gb_acq_configure": (_i32, [_vp, _i32, _f32, _i32, _vp]),
    "gb_acq_configure_once": (_i32, [_vp, _i32, _f32, _i32]),
    "gb_acq_supported_sizes": (_i32, [_vp, _i32]),
    "gb_acq_make_doppler_tables": (_i32, [_vp, _f32, _vp, _i32, _vp]),
    "gb_acq_set_doppler_tables": (_i32, [_vp, _vp, _vp, _i32]),
    "gb_acq_get_doppler_tables": (_i32, [_vp, _vp, _vp]),
    "gb_acq_set_coherent": (_i32, [_vp, _i32]),
    "gb_acq_set_mode": (_i32, [_vp, _i32]),
    "gb_acq_set_doppler_aliasing": (_i32, [_vp, _i32]),
    "gb_acq_forward_bins": (_i32, [_vp]),
    "gb_acq_set_detector": (_i32, [_vp, _f32, _i32]),
    "gb_acq_search_cells": (_i32, [_vp, _vp, _u64, _i32, _u32, _vp, _vp]),
    "gb_acq_search_cells_ring": (_i32, [_vp, _u64, _i32, _u32, _vp, _vp]),
    "gb_acq_decide": (_i32, [_vp, _vp, _i32, _i32, _i32, _f32, _u64, _f32, _vp]),
    "gb_acq_search": (_i32, [_vp, _vp, _u64, _i32, _u64, _u32, _vp, _vp]),
    "gb_acq_search_enqueue": (_i32, [_vp, _vp, _u64, _i32, _u64, _u32, _vp, _i32]),
    "gb_acq_search_wait": (_i32, [_vp, _i32, _vp, _vp]),
    "gb_acq_search_batch": (_i32, [_vp, _vp, _i32, _u64, _i32, _u64, _u32, _vp, _vp]),
    "gb_acq_search_ring": (_i32, [_vp, _u64, _i32, _u32, _vp, _vp]),
    "gb_acq_bin_power": (_i32, [_vp, _vp, _u64, _i32, _i32, _i32, _vp]),
    "gb_acq_last_kernel_ms": (_f32, [_vp]),
    "gb_acq_fine_doppler": (_i32, [_vp, _vp, _u64, _f32, _i32, _i32, _vp, _i32, _vp, _vp, _vp]),
    "gb_acq_fine_doppler_ring": (_i32, [_vp, _u64, _u64, _f32, _i32, _i32, _vp, _i32, _vp, _vp]),
    "gb_acq_fine_last_kernel_ms": (_f32, [_vp]),
    "gb_bench_fp32_tflops": (_i32, [_vp, _vp]),
    "gb_fft_c2c": (_i32, [_vp, _i32, _i32, _vp, _vp, _i32]),
    "gb_fft_power_spectrum": (_i32, [_vp, _i32, _vp, _vp, _i32]),
    "gb_rfft": (_i32, [_vp, _i32, _vp, _vp, _i32]),
    "gb_rfft_power_spectrum": (_i32, [_vp, _i32, _vp, _vp, _i32]),
    "gb_fft_c2c_f64": (_i32, [_vp, _i32, _i32, _vp, _vp, _i32]),
    "gb_fft_power_spectrum_f64": (_i32, [_vp, _i32, _vp, _vp, _i32]),
    "gb_rfft_f64": (_i32, [_vp, _i32, _vp, _vp, _i32]),
    "gb_rfft_power_spectrum_f64": (_i32, [_vp, _i32, _vp, _vp, _i32]),
    "gb_trk_channel_init": (_i32, [_vp, C.c_uint8, _f32]),
    "gb_trk_channel_start": (_i32, [_vp, _vp]),
    "gb_trk_channel_start_corrected": (_i32, [_vp, _vp]),
    "gb_trk_channel_reset": (_i32, [_vp]),
    "gb_loop_filter_new": (_i32, [_f32, _f32, _f32, _vp, _vp]),
    "gb_trk_correlate": (_i32, [_vp, _vp, _i32, _vp, _u64, _vp, _i32, _vp]),
    "gb_trk_epoch": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "gb_trk_upload": (_i32, [_vp, _vp, _i32]),
    "gb_trk_run": (_i32, [_vp, _i32, _i32, _vp]),
    "gb_trk_run_keep": (_i32, [_vp, _i32, _i32]),
    "gb_trk_download": (_i32, [_vp, _vp, _i32]),
    "gb_trk_last_kernel_ms": (_f32, [_vp]),
    "gb_nav_bit_sync": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _i32]),
    "gb_shard_prn_mask": (_u32, [_i32, _i32, _i32, _u32]),
    "gb_shard_range": (_i32, [_i32, _i32, _i32, _vp, _vp]),
    "gb_group_unique_id": (_i32, [_vp]),
    "gb_group_init": (_i32, [_vp, _vp, _i32, _i32, _vp]),
    "gb_group_rank": (_i32, [_vp]),
    "gb_group_world": (_i32, [_vp]),
    "gb_group_allgather": (_i32, [_vp, _vp, _u64, _vp]),
    "gb_group_allgather_begin": (_i32, [_vp, _vp, _u64, _i32]),
    "gb_group_allgather_end": (_i32, [_vp, _i32, _vp]),
    "gb_group_gather_results": (_i32, [_vp, _vp, _i32, _vp]),
    "gb_group_destroy": (_i32, [_vp]),
}

_LIB = None


def lib():
    """Load libgnss_b200.so; raise (never fall back) if it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libgnss_b200.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'`"
                              % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def tuning_set(key, value):
    """Explicit A/B switch of the kernels (gb_tuning_set); the library reads no environment variables."""
    check(lib().gb_tuning_set(key.encode(), int(value)), "gb_tuning_set")


def _strerror(code):
    try:
        return lib().gb_strerror(code).decode()
    except Exception:
        return "error"


def ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def check(code, what, handle=None):
    if code != GB_OK:
        detail = ""
        if handle is not None and code == GB_ECUDA:
            detail = lib().gb_last_cuda_error(handle).decode()
        raise GnssB200Error(code, what, detail)


class Handle:
    """Owns one gb_handle (one GPU)."""

    def __init__(self, device=0, ring_capacity=0):
        L = lib()
        cfg = Config(device, ring_capacity, 0)
        h = C.c_void_p()
        check(L.gb_create(C.byref(cfg), C.byref(h)), "gb_create")
        self.h = h
        self.L = L

    def close(self):
        if getattr(self, "h", None):
            self.L.gb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def call(self, name, *args):
        check(getattr(self.L, name)(self.h, *args), name, self.h)
