"""Deterministic synthetic IQ source.

The reference's sdr_mock (src/sdr_mock/device_mock.rs) is a control-plane mock that produces no
samples, and its bundled recording is missing, so the build needs its own seeded generator:
  * `if_recording`: int8 real IF samples shaped like src/test_data/GPS_recordings/config.txt:2-17
    (fs 16.3676 MHz, IF 4.1304 MHz, the ten listed PRNs with their carriers / code phases);
  * `baseband`: complex baseband GPS L1 C/A at any sample rate (BASELINE configs 2 and 3);
  * `reference_test_signal`: the noise-free waveform of do_tracking.rs:434-462, quirk included.
All phases are evaluated in f64 and rounded once.
"""
import numpy as np

# config.txt:8-17: PRN, carrier (Hz), code phase (1-based samples), strongest first
CONFIG_TXT = [(2, 4.128460e6, 15042), (3, 4.127190e6, 1618), (19, 4.129280e6, 6184), (14, 4.133130e6, 14540),
              (18, 4.127310e6, 344), (11, 4.133280e6, 2955), (32, 4.134060e6, 6857), (6, 4.127220e6, 7828),
              (28, 4.132022e6, 15203), (9, 4.132420e6, 9437)]
CONFIG_FS = 16367600.0
CONFIG_IF = 4130400.0

_G2_TAPS = [(2, 6), (3, 7), (4, 8), (5, 9), (1, 9), (2, 10), (1, 8), (2, 9), (3, 10), (2, 3), (3, 4), (5, 6), (6, 7),
            (7, 8), (8, 9), (9, 10), (1, 4), (2, 5), (3, 6), (4, 7), (5, 8), (6, 9), (1, 3), (4, 6), (5, 7), (6, 8),
            (7, 9), (8, 10), (1, 6), (2, 7), (3, 8), (4, 9)]


def ca_code(prn):
    """1023 chips, +1 for bit 1 (the convention of constants/gps_ca_constants.rs)."""
    g1 = [1] * 10
    g2 = [1] * 10
    a, b = _G2_TAPS[prn - 1]
    out = np.empty(1023, np.int8)
    for c in range(1023):
        out[c] = 1 if (g1[9] ^ g2[a - 1] ^ g2[b - 1]) else -1
        f1 = g1[2] ^ g1[9]
        f2 = g2[1] ^ g2[2] ^ g2[5] ^ g2[7] ^ g2[8] ^ g2[9]
        g1 = [f1] + g1[:9]
        g2 = [f2] + g2[:9]
    return out


def _signal(prn, n_samples, fs, carrier_hz, code_phase_samples, amp, code_doppler=0.0, nav_seed=None, phase0=0.0,
            real=False, nav_bits=None):
    t = np.arange(n_samples, dtype=np.float64)
    code = ca_code(prn).astype(np.float64)
    rate = 1.023e6 * (1.0 + code_doppler)
    # the code period starts at sample `code_phase_samples`
    chip = ((t - code_phase_samples) * rate / fs) % 1023.0
    sig = code[np.floor(chip).astype(np.int64) % 1023]
    if nav_seed is not None:
        rng = np.random.default_rng(nav_seed)
        periods = np.floor((t - code_phase_samples) * rate / fs / 1023.0).astype(np.int64)
        bits = rng.integers(0, 2, size=int(periods.max() // 20 + 3)) * 2 - 1
        sig = sig * bits[(periods // 20) + 1]
    elif nav_bits is not None:
        # a caller-planted 50 bps stream (+-1), repeated as needed; bit k covers code periods 20 k .. 20 k + 19
        periods = np.floor((t - code_phase_samples) * rate / fs / 1023.0).astype(np.int64)
        nb = np.asarray(nav_bits, np.float64)
        sig = sig * nb[((periods // 20) + 1) % len(nb)]
    ph = 2.0 * np.pi * ((carrier_hz * t / fs) % 1.0) + phase0
    if real:
        return amp * sig * np.cos(ph)
    return amp * sig * np.exp(1j * ph)


def if_recording(n_ms=10, seed=0x6E55, noise_sigma=8.0, amp0=4.0, amp_step=0.9, prns=None):
    """int8 real IF stand-in for gioveAandB_short.bin."""
    n = int(round(CONFIG_FS / 1000.0)) * n_ms
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(n) * noise_sigma
    truth = []
    for k, (prn, carr, phase1) in enumerate(CONFIG_TXT):
        if prns is not None and prn not in prns:
            continue
        amp = amp0 * (amp_step ** k)
        x += _signal(prn, n, CONFIG_FS, carr, phase1 - 1, amp, phase0=rng.uniform(0, 2 * np.pi), real=True)
        truth.append({"prn": prn, "carrier": carr, "code_phase": phase1 - 1, "amp": amp})
    return np.clip(np.round(x), -128, 127).astype(np.int8), truth


def i8_to_c32(x):
    """do_acquisition.rs:420-424: Complex32::new(x as i8 as f32, 0.0)."""
    return x.astype(np.float32).astype(np.complex64)


def baseband(fs, n_ms, sats, seed=0x6E56, noise_sigma=1.0, nav=False):
    """Complex baseband.  sats: list of dicts {prn, doppler, code_phase (samples), cn0_dbhz}."""
    n = int(round(fs / 1000.0)) * n_ms
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * (noise_sigma / np.sqrt(2.0))
    for k, s in enumerate(sats):
        # C/N0 = A^2 / (sigma^2 / fs)  =>  A = sigma * sqrt(10^(cn0/10) / fs)
        amp = noise_sigma * np.sqrt(10.0 ** (s["cn0_dbhz"] / 10.0) / fs)
        x += _signal(s["prn"], n, fs, s["doppler"], s["code_phase"], amp, code_doppler=s["doppler"] / 1575.42e6,
                     nav_seed=(seed * 131 + k) if (nav and s.get("nav_bits") is None) else None,
                     phase0=rng.uniform(0, 2 * np.pi), nav_bits=s.get("nav_bits"))
    return x.astype(np.complex64)


def reference_test_signal(code_samples, doppler, carrier_phase0, code_phase0, fs):
    """generate_synthetic_signal (do_tracking.rs:434-462) in f32, including its quirk of indexing the
    already-resampled code with a chip index."""
    n = int(np.float32(fs) / np.float32(1000.0))
    i = np.arange(n, dtype=np.float32)
    step = np.float32(1.023e6) / np.float32(fs)
    cph = np.float32(carrier_phase0) + (np.float32(2.0) * np.float32(np.pi) * np.float32(doppler) / np.float32(fs) * i)
    cur = np.float32(code_phase0) + step * i
    idx = np.floor(cur).astype(np.int64) % 1023
    cv = np.asarray(code_samples)[idx].astype(np.float32)
    return (cv * np.cos(cph).astype(np.float32) + 1j * (cv * np.sin(cph).astype(np.float32))).astype(np.complex64)


# ---------------------------------------------------------------------------------------------
# Other constellations (BASELINE configs 4-5).  NOTHING here comes from the reference (GPS L1 C/A only);
# values are recalled from the public ICDs and cannot be verified offline.  They are
# used consistently by the generator and the correlator, so they only affect synthetic self-consistency.
_B1I_TAPS = [(1, 3), (1, 4), (1, 5), (1, 6), (1, 8), (1, 9), (1, 10), (1, 11), (2, 7), (3, 4), (3, 5), (3, 6), (3, 8),
             (3, 9), (3, 10), (3, 11), (4, 5), (4, 6), (4, 8), (4, 9), (4, 10), (4, 11), (5, 6), (5, 8), (5, 9), (5, 10),
             (5, 11), (6, 8), (6, 9), (6, 10), (6, 11), (8, 9), (8, 10), (8, 11), (9, 10), (9, 11), (10, 11)]


def b1i_code(prn):
    """BeiDou B1I ranging code: 2046 chips of the truncated 11-stage Gold sequence, +1 for bit 1.
    G1 = x^11+x^10+x^9+x^8+x^7+x+1, G2 = x^11+x^9+x^8+x^5+x^4+x^3+x^2+x+1, both initialised 01010101010."""
    g1 = [0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0]
    g2 = list(g1)
    a, b = _B1I_TAPS[prn - 1]
    out = np.empty(2046, np.int8)
    for c in range(2046):
        out[c] = 1 if (g1[10] ^ g2[a - 1] ^ g2[b - 1]) else -1
        f1 = g1[0] ^ g1[6] ^ g1[7] ^ g1[8] ^ g1[9] ^ g1[10]
        f2 = g2[0] ^ g2[1] ^ g2[2] ^ g2[3] ^ g2[4] ^ g2[7] ^ g2[8] ^ g2[10]
        g1 = [f1] + g1[:10]
        g2 = [f2] + g2[:10]
    return out


def e1_surrogate_code(prn, seed=0xE1C0DE):
    """Galileo E1 primary codes are ICD memory codes (not LFSR-generable, not available offline): a clearly
    labelled deterministic SURROGATE of 4092 +-1 chips per PRN from a seeded PRNG."""
    rng = np.random.default_rng(seed + 7919 * prn)
    return (rng.integers(0, 2, 4092) * 2 - 1).astype(np.int8)


def resample_code(chips, chip_rate, fs, n_samples, boc11=False):
    """Code samples at fs over n_samples (one code period): chip index = floor(x * chip_rate / fs) (the f32 index
    arithmetic of ca_code.rs:18-22), optionally times the BOC(1,1) sub-carrier sign(sin(2 pi chip_phase))."""
    x = np.arange(n_samples, dtype=np.float32)
    ph = (x * np.float32(chip_rate)) / np.float32(fs)
    idx = np.floor(ph).astype(np.int64) % len(chips)
    s = np.asarray(chips)[idx].astype(np.int8)
    if boc11:
        s = s * np.where((ph - np.floor(ph)) < 0.5, 1, -1).astype(np.int8)
    return s


def multi_gnss(fs, n_ms, sats, seed=0x6E58, noise_sigma=1.0):
    """Complex baseband with GPS L1 C/A ('G'), BeiDou B1I ('C') and Galileo-E1-like BOC(1,1) ('E') signals.
    sats: dicts {system, prn, doppler, code_phase (samples), cn0_dbhz}."""
    n = int(round(fs / 1000.0)) * n_ms
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * (noise_sigma / np.sqrt(2.0))
    t = np.arange(n, dtype=np.float64)
    for s in sats:
        sysname = s["system"]
        chips, rate = {"G": (ca_code(s["prn"]), 1.023e6), "C": (b1i_code(s["prn"]), 2.046e6),
                       "E": (e1_surrogate_code(s["prn"]), 1.023e6)}[sysname]
        ph = ((t - s["code_phase"]) * rate / fs) % len(chips)
        sig = chips[np.floor(ph).astype(np.int64) % len(chips)].astype(np.float64)
        if sysname == "E":
            sig = sig * np.where((ph - np.floor(ph)) < 0.5, 1.0, -1.0)
        amp = noise_sigma * np.sqrt(10.0 ** (s["cn0_dbhz"] / 10.0) / fs)
        x += amp * sig * np.exp(1j * (2.0 * np.pi * ((s["doppler"] * t / fs) % 1.0) + rng.uniform(0, 2 * np.pi)))
    return x.astype(np.complex64)
