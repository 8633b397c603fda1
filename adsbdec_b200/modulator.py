"""Synthetic Mode S captures in the format the reference's ``-f`` path reads.

The reference reads little-endian ``unsigned short`` samples centred on 0x800
(airspy R2 ``AIRSPY_SAMPLE_RAW``: 12-bit, real-valued, 20 MS/s, signal at fs/4;
/root/reference/air.c:30, :54, :64, :120 - SURVEY.md F1).  The waveform follows
SURVEY.md App. E: 0.5 us pulses (10 samples), preamble pulses at 0 / 1.0 / 3.5 /
4.5 us, data from 8.0 us at 1 us per bit, MSB first, '1' = pulse in the first
half; carrier ``cos(pi/2 * t + phi)``; additive white Gaussian noise; 12-bit clip.

Nothing here is on the decode path; it only feeds tests and ``bench.py``.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

FS = 20_000_000            # input sample rate (air.c:30)
PULSE = 10                 # input samples per 0.5 us pulse
MID = 2048                 # mid-scale of the 12-bit ADC (air.c:64)
GEN = 0xFFF409             # Mode S CRC-24 generator (crc.h table == this polynomial)
PAD_SAMPLES = 1 << 21      # mid-scale tail so the reference drains its queue


def crc24(data: bytes) -> int:
    """CRC-24 remainder of ``data`` (bitwise, poly 0xFFF409)."""
    r = 0
    for byte in data:
        r ^= byte << 16
        for _ in range(8):
            r = ((r << 1) ^ GEN) & 0xFFFFFF if r & 0x800000 else (r << 1) & 0xFFFFFF
    return r


def make_frame(df: int, rng: np.random.Generator, *, ca: int | None = None,
               overlay: int = 0, aa: int | None = None) -> bytes:
    """A 56-bit (DF<16) or 112-bit frame with valid parity.

    ``overlay`` is XORed onto the parity field: 0 gives the syndrome-0 frames the
    reference accepts (valid.c:52, :74); an ICAO address gives DF20/21-style
    address/parity, an interrogator id gives DF11 with II != 0.
    """
    nbytes = 7 if df < 16 else 14
    body = bytearray(rng.integers(0, 256, nbytes - 3, dtype=np.uint8).tobytes())
    body[0] = ((df & 0x1F) << 3) | ((int(rng.integers(0, 8)) if ca is None else ca) & 7)
    if aa is not None:          # the AA field of DF11 / DF17 / DF18: bytes 1..3
        body[1:4] = int(aa).to_bytes(3, "big")
    par = crc24(bytes(body)) ^ overlay
    return bytes(body) + par.to_bytes(3, "big")


def flip_bits(frame: bytes, positions) -> bytes:
    """Flip the given bit positions (0 = MSB of byte 0)."""
    b = bytearray(frame)
    for p in positions:
        b[p // 8] ^= 0x80 >> (p % 8)
    return bytes(b)


def envelope(frame: bytes) -> np.ndarray:
    """0/1 pulse envelope of one frame at 20 MS/s (preamble + PPM data)."""
    nbits = 8 * len(frame)
    env = np.zeros(16 * PULSE + nbits * 2 * PULSE, dtype=np.float32)
    for k in (0, 2, 7, 9):                      # 0, 1.0, 3.5, 4.5 us
        env[k * PULSE:(k + 1) * PULSE] = 1.0
    bits = np.unpackbits(np.frombuffer(frame, dtype=np.uint8))
    base = 16 * PULSE
    for i, bit in enumerate(bits):
        s = base + i * 2 * PULSE + (0 if bit else PULSE)
        env[s:s + PULSE] = 1.0
    return env


@dataclass
class Burst:
    start: int          # first input sample of the preamble
    frame: bytes        # bytes as transmitted (after any injected bit errors)
    amplitude: float    # carrier amplitude in LSB
    phase: float        # carrier phase at the first sample


@dataclass
class Capture:
    samples: np.ndarray                 # uint16, little-endian on disk
    bursts: list[Burst] = field(default_factory=list)
    seed: int = 0

    def tofile(self, path) -> None:
        self.samples.astype("<u2").tofile(path)


def render(n_samples: int, bursts: list[Burst], sigma: float, rng: np.random.Generator,
           pad: int = PAD_SAMPLES) -> np.ndarray:
    """Sum the bursts, add AWGN, quantise to 12 bits, append the mid-scale pad."""
    n_samples -= n_samples % 8              # byte length 0 mod 8 (SURVEY 8a fileInput row)
    sig = np.zeros(n_samples, dtype=np.float32)
    for b in bursts:
        env = envelope(b.frame)
        end = min(b.start + env.size, n_samples)
        if b.start >= n_samples or end <= b.start:
            continue
        t = np.arange(b.start, end, dtype=np.float64)
        car = np.cos(0.5 * np.pi * t + b.phase).astype(np.float32)
        sig[b.start:end] += np.float32(b.amplitude) * env[:end - b.start] * car
    if sigma > 0:
        sig += rng.standard_normal(n_samples, dtype=np.float32) * np.float32(sigma)
    x = np.clip(np.rint(sig + MID), 0, 4095).astype("<u2")
    if pad:
        x = np.concatenate([x, np.full(pad, MID, dtype="<u2")])
    return x


def amplitude_for_snr(snr_db: float, sigma: float) -> float:
    """Carrier amplitude with pulse power A^2/2 = sigma^2 * 10^(snr/10)."""
    return float(sigma * np.sqrt(2.0 * 10.0 ** (snr_db / 10.0)))


def synth_capture(seconds: float, n_frames: int, *, snr_db: float = 20.0, sigma: float = 25.0,
                  seed: int = 0, df_mix=((17, 1.0),), bit_errors: int = 0,
                  overlay_for=None, pad: int = PAD_SAMPLES, poisson: bool = True) -> Capture:
    """Random traffic: ``n_frames`` bursts uniformly (Poisson) placed, overlaps allowed.

    ``df_mix``: (df, weight) pairs.  ``bit_errors``: 0, or flip that many random
    bits in every other frame.  ``overlay_for``: dict df -> True to overlay a random
    24-bit address on the parity (DF20/21/4/5 style).
    """
    rng = np.random.default_rng(seed)
    n_samples = int(round(seconds * FS))
    n_samples -= n_samples % 8
    dfs = np.array([d for d, _ in df_mix])
    w = np.array([x for _, x in df_mix], dtype=np.float64)
    w /= w.sum()
    amp = amplitude_for_snr(snr_db, sigma)
    if poisson:
        starts = np.sort(rng.integers(0, max(1, n_samples - 2600), n_frames))
    else:
        starts = (np.arange(n_frames) * ((n_samples - 2600) // max(1, n_frames))).astype(np.int64)
    bursts = []
    for i, s in enumerate(starts):
        df = int(rng.choice(dfs, p=w))
        ov = int(rng.integers(1, 1 << 24)) if overlay_for and overlay_for.get(df) else 0
        fr = make_frame(df, rng, overlay=ov)
        if bit_errors and i % 2 == 1:
            fr = flip_bits(fr, rng.choice(8 * len(fr), bit_errors, replace=False))
        bursts.append(Burst(int(s), fr, amp * float(rng.uniform(0.7, 1.4)),
                            float(rng.uniform(0, 2 * np.pi))))
    return Capture(render(n_samples, bursts, sigma, rng, pad=pad), bursts, seed)


def synth_fleet_capture(seconds: float, n_frames: int, n_aircraft: int, *, snr_db: float = 20.0, sigma: float = 25.0,
                        seed: int = 0, strangers: float = 0.15, pad: int = PAD_SAMPLES) -> Capture:
    """Traffic of a fleet: every aircraft has an ICAO address, announces it in DF11 / DF17 / DF18 frames (AA field,
    plain parity) and answers interrogations with DF4 / DF5 / DF20 / DF21 frames whose parity field is overlaid
    with that address.  A share ``strangers`` of the address/parity frames comes from aircraft that never announce
    themselves.  For the address/parity extension; ``Burst.frame`` keeps what was sent."""
    rng = np.random.default_rng(seed)
    n_samples = int(round(seconds * FS))
    n_samples -= n_samples % 8
    fleet = rng.choice(np.arange(1, 1 << 24), n_aircraft, replace=False)
    amp = amplitude_for_snr(snr_db, sigma)
    starts = np.sort(rng.integers(0, max(1, n_samples - 2600), n_frames))
    bursts = []
    for s in starts:
        icao = int(rng.choice(fleet))
        kind = rng.random()
        if kind < 0.45:
            fr = make_frame(int(rng.choice([11, 17, 17, 18])), rng, aa=icao, ca=(0 if kind < 0.1 else None))
        else:
            if rng.random() < strangers:
                icao = int(rng.integers(1, 1 << 24))
            fr = make_frame(int(rng.choice([4, 5, 20, 21])), rng, overlay=icao)
        bursts.append(Burst(int(s), fr, amp * float(rng.uniform(0.7, 1.4)), float(rng.uniform(0, 2 * np.pi))))
    return Capture(render(n_samples, bursts, sigma, rng, pad=pad), bursts, seed)


def synth_iq_capture(seconds: float, n_frames: int, *, snr_db: float = 20.0, sigma: float = 40.0, seed: int = 0,
                     df_mix=((17, 1.0),), pad: int = PAD_SAMPLES, fleet: int = 0) -> Capture:
    """The same traffic as complex baseband: signed 16-bit I/Q pairs at 10 MS/s (what libairspy hands out as
    INT16_IQ: it does the fs/4 translation and the decimation itself).  ``samples`` holds the interleaved pairs
    reinterpreted as uint16 - one pair per envelope sample - ready for ``Decoder(iq=True).push``.  ``fleet`` > 0:
    frames of that many aircraft as in synth_fleet_capture (for the address/parity extension).  ``Burst.start`` is in
    16-bit values, like everywhere else (two per envelope sample)."""
    rng = np.random.default_rng(seed)
    n_pairs = int(round(seconds * FS / 2))
    n_pairs -= n_pairs % 4
    dfs = np.array([d for d, _ in df_mix])
    w = np.array([x for _, x in df_mix], dtype=np.float64)
    w /= w.sum()
    amp = amplitude_for_snr(snr_db, sigma)
    aircraft = rng.choice(np.arange(1, 1 << 24), fleet, replace=False) if fleet else None
    starts = np.sort(rng.integers(0, max(1, n_pairs - 1300), n_frames))
    z = np.zeros(n_pairs, dtype=np.complex64)
    bursts = []
    for s in starts:
        if fleet:
            icao = int(rng.choice(aircraft))
            if rng.random() < 0.45:
                fr = make_frame(int(rng.choice([11, 17, 17, 18])), rng, aa=icao, ca=(0 if rng.random() < 0.2 else None))
            else:
                fr = make_frame(int(rng.choice([4, 5, 20, 21])), rng, overlay=icao)
        else:
            fr = make_frame(int(rng.choice(dfs, p=w)), rng)
        a = amp * float(rng.uniform(0.7, 1.4))
        ph = float(rng.uniform(0, 2 * np.pi))
        env = envelope(fr)[::2]                 # 20 MS/s pulse shape -> 10 MS/s (5 samples per half microsecond)
        end = min(int(s) + env.size, n_pairs)
        z[int(s):end] += np.complex64(a * np.exp(1j * ph)) * env[:end - int(s)]
        bursts.append(Burst(2 * int(s), fr, a, ph))
    if sigma > 0:
        z += (rng.standard_normal(n_pairs, dtype=np.float32) + 1j * rng.standard_normal(n_pairs, dtype=np.float32)) \
            * np.float32(sigma)
    iq = np.empty(2 * n_pairs, dtype=np.int16)
    iq[0::2] = np.clip(np.rint(z.real), -8191, 8191)
    iq[1::2] = np.clip(np.rint(z.imag), -8191, 8191)
    if pad:
        iq = np.concatenate([iq, np.zeros(pad, dtype=np.int16)])
    return Capture(iq.view("<u2"), bursts, seed)
