"""Host-side mirror of src/tracking/do_tracking.rs over libgnss_b200: TrackingChannel state (pub
fields, :88-115), start/reset (:148-154, :311-326), and the batched replacements of
early_late_correlation (:231-272), do_work (:183-210) and TrackingManager::process_channels (:350-372)."""
import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import GB_TRK_FAST, GB_TRK_ORDERED, GB_TRK_IDLE, GB_TRK_TRACKING, TrkChannel  # noqa: F401

NUM_OF_CHANNELS = 15  # do_tracking.rs:18


def channel_array(n, fs):
    """n x TrackingChannel::new(id, fs)."""
    arr = (TrkChannel * n)()
    L = _ffi.lib()
    for i in range(n):
        _ffi.check(L.gb_trk_channel_init(C.byref(arr[i]), i & 0xFF, float(fs)), "gb_trk_channel_init")
    return arr


def start(ch, prn, carrier_freq, code_phase_chips, sample_global_index, fs, code_row=None, corrected=False):
    """TrackingChannel::start(AcquisitionResult).  By default the reference's row (= prn, Q6: the channel correlates
    with PRN + 1's code); corrected=True uses the satellite's own row (prn - 1, gb_trk_channel_start_corrected);
    code_row overrides both."""
    r = _ffi.AcqResult()
    r.prn, r.found = int(prn), 1
    r.carrier_freq, r.code_phase_chips, r.fs = float(carrier_freq), float(code_phase_chips), float(fs)
    r.sample_global_index = int(sample_global_index)
    fn = "gb_trk_channel_start_corrected" if corrected else "gb_trk_channel_start"
    _ffi.check(getattr(_ffi.lib(), fn)(C.byref(ch), C.byref(r)), fn)
    if code_row is not None:
        ch.code_row = int(code_row)


def loop_filter(noise_bw, damping, gain):
    t1, t2 = C.c_float(), C.c_float()
    _ffi.check(_ffi.lib().gb_loop_filter_new(float(noise_bw), float(damping), float(gain), C.byref(t1), C.byref(t2)),
               "gb_loop_filter_new")
    return t1.value, t2.value


class TrackingEngine:
    """All tracking channels of one GPU."""

    def __init__(self, handle):
        self.hd = handle

    def correlate(self, channels, data_list, mode=GB_TRK_FAST):
        """early_late_correlation for each channel on its own samples (data_list[c])."""
        n = len(channels)
        offs = np.zeros(n, np.uint64)
        total = 0
        for c in range(n):
            offs[c] = total
            total += len(data_list[c])
        data = np.concatenate([np.ascontiguousarray(d, np.complex64) for d in data_list])
        out = np.zeros(n, _ffi.CORR_DTYPE)
        # the library checks every segment against the buffer length (a short segment -> GB_ERANGE)
        self.hd.call("gb_trk_correlate", channels, n, _ffi.ptr(data), int(data.size), _ffi.ptr(offs), int(mode),
                     _ffi.ptr(out))
        return out

    def epoch(self, channels, mode=GB_TRK_FAST):
        """One do_work per active channel whose samples are in the ring."""
        n = len(channels)
        out = np.zeros(n, _ffi.CORR_DTYPE)
        ran = np.zeros(n, np.uint8)
        lost = np.zeros(n, np.uint8)
        self.hd.call("gb_trk_epoch", channels, n, int(mode), _ffi.ptr(out), _ffi.ptr(ran), _ffi.ptr(lost))
        return out, ran, lost

    def upload(self, channels):
        self.n = len(channels)
        self.hd.call("gb_trk_upload", channels, self.n)

    def run(self, n_epochs, mode=GB_TRK_FAST, want_hist=False, keep_on_device=False):
        """n_epochs do_work epochs per channel in one launch.  keep_on_device: the prompt history stays in HBM for
        nav_bit_sync(handle, None, ...) instead of coming back to the host."""
        if keep_on_device:
            self.hd.call("gb_trk_run_keep", int(n_epochs), int(mode))
            self.kept = (int(n_epochs), self.n)
            return None
        hist = np.zeros((n_epochs, self.n, 2), np.float32) if want_hist else None
        self.hd.call("gb_trk_run", int(n_epochs), int(mode), _ffi.ptr(hist))
        return hist

    def download(self, channels):
        self.hd.call("gb_trk_download", channels, len(channels))
        return channels

    def last_kernel_ms(self):
        return float(self.hd.L.gb_trk_last_kernel_ms(self.hd.h))


def nav_bit_sync(handle, prompt_hist, max_bits=4096, shape=None):
    """Bit sync + 20 ms prompt accumulation + preamble search (N4) on a prompt history [n_epochs, n_channels, 2];
    prompt_hist=None with shape=(n_epochs, n_channels): the history TrackingEngine.run(keep_on_device=True) left in HBM."""
    if prompt_hist is None:
        hist, (n_epochs, n_channels) = None, shape
    else:
        hist = np.ascontiguousarray(prompt_hist, np.float32)
        n_epochs, n_channels = hist.shape[0], hist.shape[1]
    st = np.zeros(n_channels, _ffi.NAV_DTYPE)
    bits = np.zeros((n_channels, max_bits), np.int8)
    handle.call("gb_nav_bit_sync", _ffi.ptr(hist), int(n_epochs), int(n_channels), _ffi.ptr(st), _ffi.ptr(bits), int(max_bits))
    return st, bits
