"""The C-ABI library: loads, exports every symbol include/gnss_b200.h declares, host-only entry points agree with
the oracle, and compute entry points fail loudly without a device (no CPU fallback).  CPU only."""
import ctypes as C
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "gnss_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(gb_[a-z0-9_]+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported_and_bound(ffi):
    names = _declared()
    assert len(names) >= 40
    out = subprocess.run(["nm", "-D", "--defined-only", ffi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (gb_[a-z0-9_]+)", out))
    assert set(names) <= exported, sorted(set(names) - exported)
    assert set(names) == set(ffi.SIGNATURES), sorted(set(names) ^ set(ffi.SIGNATURES))
    L = ffi.lib()
    for n in names:
        assert getattr(L, n) is not None


def test_library_is_sm100a_native_cuda(ffi):
    out = subprocess.run(["cuobjdump", "-lelf", ffi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "gnss_b200.h"\nint main(void){gb_acq_cell c; gb_trk_channel t; (void)c; (void)t; return sizeof(gb_acq_result) > 0 ? 0 : 1;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                    "-o", str(tmp_path / "t.o")], check=True)


def test_struct_layouts_match_header(ffi, tmp_path):
    src = tmp_path / "s.c"
    src.write_text('#include <stdio.h>\n#include "gnss_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(gb_acq_cell),'
                   ' sizeof(gb_acq_result), sizeof(gb_trk_channel), sizeof(gb_trk_corr), sizeof(gb_config), sizeof(gb_fine_req),'
                   ' sizeof(gb_fine_result), sizeof(gb_nav_sync));return 0;}\n')
    exe = tmp_path / "s"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [ffi.CELL_DTYPE.itemsize, C.sizeof(ffi.AcqResult), C.sizeof(ffi.TrkChannel), ffi.CORR_DTYPE.itemsize,
                     C.sizeof(ffi.Config), ffi.FINE_REQ_DTYPE.itemsize, ffi.FINE_RES_DTYPE.itemsize, ffi.NAV_DTYPE.itemsize]


def test_no_device_means_error_not_fallback(ffi):
    L = ffi.lib()
    if L.gb_device_count() > 0:
        return  # on the GPU box the gpu-marked tests cover the device path
    h = C.c_void_p()
    assert L.gb_create(None, C.byref(h)) == ffi.GB_ENODEVICE
    assert h.value is None
    try:
        ffi.Handle(0)
        raise AssertionError("Handle() must raise without a device")
    except ffi.GnssB200Error as e:
        assert e.code == ffi.GB_ENODEVICE


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gnss-sdr-rs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".sh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "gnss_oracle" not in txt and "libgnss_oracle" not in txt, f


def test_host_helpers_match_oracle(ffi, oracle):
    L, O = ffi.lib(), oracle.lib()
    tab = oracle.ca_table()
    for prn in range(1, 33):
        out = np.zeros(1023, np.int8)
        assert L.gb_ca_code_chips(prn, ffi.ptr(out)) == 0
        assert (out == tab[prn - 1]).all()
    assert L.gb_ca_code_chips(33, ffi.ptr(np.zeros(1023, np.int8))) == ffi.GB_EINVAL
    for fs in (2.048e6, 4.092e6, 4.096e6, 16367600.0, 20e6):
        n = L.gb_num_samples_per_code(1.023e6, fs)
        assert n == O.go_num_samples_per_code(1.023e6, fs)
        a = np.zeros(n, np.int8)
        assert L.gb_generate_ca_code_samples(19, 1.023e6, fs, ffi.ptr(a), n) == n
        assert (a == oracle.ca_code_samples(19, 1.023e6, fs)).all()
    # loop filter + channel init / start / reset
    t1, t2 = C.c_float(), C.c_float()
    for args in ((25.0, 0.7, 0.25), (2.0, 0.7, 1.0)):
        L.gb_loop_filter_new(*args, C.byref(t1), C.byref(t2))
        ref = O.go_loop_filter_new(*args)
        assert (t1.value, t2.value) == (ref.tau1, ref.tau2)
    ch = ffi.TrkChannel()
    L.gb_trk_channel_init(C.byref(ch), 4, 4.096e6)
    och = oracle.trk_channel(4, 4.096e6)
    for k in ("id", "prn", "state", "lost_counter", "fs", "next_sample_index", "num_samples_per_code", "code_rate"):
        assert getattr(ch, k) == getattr(och, k), k
    assert (ch.pll_tau1, ch.pll_tau2, ch.dll_tau1, ch.dll_tau2) == (och.pll_filter.tau1, och.pll_filter.tau2,
                                                                     och.dll_filter.tau1, och.dll_filter.tau2)


def test_decide_matches_oracle_on_random_cells(ffi, oracle):
    """gb_acq_decide (host O(D) scan, Q1) against the oracle's go_acq_decide on adversarial cell tables."""
    from gnss_sdr_rs_b200 import acquisition
    rng = np.random.default_rng(5)
    n, fs = 4092, 4.092e6
    for trial in range(300):
        D = int(rng.integers(1, 40))
        cells = np.zeros(D, ffi.CELL_DTYPE)
        base = rng.uniform(0.5, 2.0, D).astype(np.float32)
        cells["sum8"] = base * n
        cells["peak"] = base * rng.choice([3.0, 6.9, 7.0, 7.2, 9.0, 30.0], D).astype(np.float32)
        if trial % 7 == 0:
            cells["peak"][:] = 0  # nothing above zero: 0/0 = NaN -> None
            cells["sum8"][:] = 0
        cells["argmax"] = rng.integers(0, n, D)
        carr = (np.arange(D) * 500.0 - 7000.0).astype(np.float32)
        got = acquisition.decide(cells, carr, 9, n, fs, local_tail=1000)
        ocells = np.zeros(D, oracle.CELL_DTYPE)
        for k in ("peak", "argmax", "sum8"):
            ocells[k] = cells[k]
        ref = oracle.acq_decide(ocells, carr, 9, n, fs, local_tail=1000)
        assert (got is None) == (ref is None)
        if ref:
            assert got["doppler_bin"] == ref["bin"]
            for k in ("code_phase_samples", "code_phase_chips", "carrier_freq", "mag_relative", "sample_global_index"):
                assert got[k] == ref[k], k


def test_sharding_helpers_match_the_python_partition(ffi):
    """gb_shard_prn_mask / gb_shard_range (host-only, no device): the partition the multi-GPU paths use, identical to
    gnss_sdr_rs_b200.sharding (which the world-size-2 gloo test exercises); every item is owned by exactly one rank."""
    from gnss_sdr_rs_b200 import sharding
    L = ffi.lib()
    for world in (1, 2, 3, 4, 8):
        for base in (0xFFFFFFFF, 0xFFFFFFFF & ~(1 << 4), 0x0000F0F1):
            seen = 0
            for rank in range(world):
                m = int(L.gb_shard_prn_mask(rank, world, 32, base))
                assert m == sharding.prn_mask_for_rank(rank, world, 32, base)
                assert seen & m == 0
                seen |= m
            assert seen == base
        for n_items in (0, 1, 13, 512, 1024):
            covered = []
            for rank in range(world):
                first, count = C.c_int(-1), C.c_int(-1)
                assert L.gb_shard_range(n_items, rank, world, C.byref(first), C.byref(count)) == 0
                assert list(range(first.value, first.value + count.value)) == list(sharding.items_for_rank(n_items, rank, world))
                covered += list(range(first.value, first.value + count.value))
            assert covered == list(range(n_items))
    assert L.gb_shard_prn_mask(2, 2, 32, 0xFFFFFFFF) == 0          # rank out of range
    f, c = C.c_int(), C.c_int()
    assert L.gb_shard_range(10, 3, 2, C.byref(f), C.byref(c)) == ffi.GB_EINVAL
    # without a GPU the collective cannot be set up, and says so (no silent fallback)
    if L.gb_device_count() == 0:
        g = C.c_void_p()
        ids = (C.c_uint8 * 128)()
        assert L.gb_group_init(None, ids, 0, 1, C.byref(g)) == ffi.GB_EINVAL


def test_acquisition_manager_mirror():
    """do_acquisition.rs:339-395 on the host mirror."""
    from gnss_sdr_rs_b200.acquisition import AcquisitionManager
    m = AcquisitionManager()
    assert m.mode == m.COLD
    assert m.get_pacing_and_list(set()) == (500, 0xFFFFFFFF)
    m.update_mode(3)
    assert m.mode == m.WARM
    assert m.get_pacing_and_list({1, 2, 3}) == (1000, 2040)
    m.update_mode(5)
    assert m.mode == m.STEADY
    m.update_mode(0)
    assert m.mode == m.COLD


def test_frontend_phase_orbit_reproduces_the_sequential_accumulator(ffi, oracle):
    """gb_frontend_orbit (host-only): the tail + cycle of `acc = (acc + step) % 2048.0` (rf/frontend.rs:48-52) found at
    configure time yields, sample for sample, the LUT indices of the oracle's sequential accumulator -- also past the
    point where the table wraps (lambda = 4, 512, 2048) and for a negative IF (saturating cast -> index 0)."""
    L = ffi.lib()
    n = 40000
    for f_if, fs, want in ((4130400.0, 16367600.0, (3, 6313323)), (4092000.0, 16368000.0, (0, 4)),
                           (420000.0, 2048000.0, (0, 512)), (1000.0, 2048000.0, (0, 2048)),
                           (-4130400.0, 16367600.0, None), (38400.0, 2048000.0, None)):
        mu, lam = C.c_uint64(), C.c_uint64()
        idx = np.zeros(n, np.uint16)
        assert L.gb_frontend_orbit(f_if, fs, C.byref(mu), C.byref(lam), idx.ctypes.data_as(C.c_void_p), n) == 0
        if want:
            assert (mu.value, lam.value) == want
        assert 1 <= lam.value <= 1 << 24
        f = oracle.frontend(f_if, fs)
        # the oracle mixes (1, 0) samples with zero DC state only if alpha were 0; read the index sequence off its
        # phase accumulator instead: one process_block of 8 samples advances it 8 steps
        ref = np.zeros(n, np.uint16)
        acc = np.float32(0.0)
        step = np.float32(np.float32(f_if) / np.float32(fs)) * np.float32(2048.0)
        assert step == np.float32(f.phase_step)
        for i in range(n):
            ref[i] = 0 if not acc > 0 else int(acc) % 2048
            acc = np.float32(np.fmod(np.float32(acc + step), np.float32(2048.0)))
        assert (idx == ref).all()
        z = np.zeros(n - n % 8, np.complex64)
        oracle.frontend_process(f, z)
        assert np.float32(f.phase_accumulator) == np.float32(_phase_after(step, n - n % 8))
    assert L.gb_frontend_orbit(1.0, 0.0, C.byref(mu), C.byref(lam), None, 0) == ffi.GB_EINVAL


def _phase_after(step, n):
    acc = np.float32(0.0)
    for _ in range(n):
        acc = np.float32(np.fmod(np.float32(acc + step), np.float32(2048.0)))
    return acc


def test_nested_radix31_butterfly_is_generated_and_correct(tmp_path):
    """dft31_nested.cuh (the 31-point butterfly of the N = 4092 plan: two 15-point cyclic correlations nested 3 x 5) is the
    output of gen_dft31_nested.py, and the generator's own checks hold: the nested correlation equals its definition, the
    butterfly equals numpy's FFT in f64 for both signs, and a float32 emulation of the emitted operation order is within
    1e-6 of the f64 DFT."""
    import re
    import subprocess
    import sys
    csrc = os.path.join(ROOT, "gnss-sdr-rs_b200", "csrc")
    out = tmp_path / "dft31_nested.cuh"
    r = subprocess.run([sys.executable, os.path.join(csrc, "gen_dft31_nested.py"), str(out)], capture_output=True, text=True, check=True)
    assert out.read_text() == open(os.path.join(csrc, "dft31_nested.cuh")).read()
    errs = [float(x) for x in re.findall(r"err ([0-9.e+-]+)", r.stdout)]
    assert len(errs) == 5 and max(errs[:3]) < 1e-12 and max(errs[3:]) < 1e-6, r.stdout

