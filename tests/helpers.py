"""Test-only helpers: capture recipes and a numpy all-offsets candidate generator.

The numpy generator restates the per-offset half of the path (demod.c:102-107 preamble test,
demod.c:46-81 DF gate, demod.c:31-44 slicer, valid.c:39-82 syndrome) on top of the ORACLE's
envelope, so that the host-side resolver can be exercised on a CPU-only box.  It is a test
double, never shipped.
"""
from __future__ import annotations

import numpy as np

from adsbdec_b200 import CAND_DTYPE
from adsbdec_b200.modulator import synth_capture

DF_MIX_ALL = ((17, 0.5), (11, 0.25), (18, 0.15), (20, 0.05), (21, 0.05))


def capture_basic(seed=1, seconds=0.2, n_frames=60, snr_db=20.0, pad=1 << 16, **kw):
    return synth_capture(seconds, n_frames, snr_db=snr_db, seed=seed, df_mix=DF_MIX_ALL, pad=pad, **kw)


def frames_key(frames):
    return [(f.n, f.ts, f.pw, f.data) for f in frames]


def crc_table():
    t = np.zeros(256, dtype=np.uint32)
    for b in range(256):
        r = b << 16
        for _ in range(8):
            r = ((r << 1) ^ 0xFFF409) if r & 0x800000 else (r << 1)
        t[b] = r & 0xFFFFFF
    return t


def all_offset_candidates(amp: np.ndarray, accept_df18: bool, n_end: int | None = None):
    """Every offset n < n_end that passes preamble + DF gate ("tries": (n, cls)) and, of those,
    the ones with syndrome 0 (structured array in the C ABI's candidate layout)."""
    a = amp.astype(np.float32, copy=False)
    N = a.size
    n_end = N - 1200 if n_end is None else min(n_end, N - 1200)
    if n_end <= 0:
        return np.zeros(0, dtype=CAND_DTYPE), np.zeros((0, 2), dtype=np.int64)
    c = (a[:-10] + a[10:]).astype(np.float32)        # a[n] + a[n+10]
    ci = c.astype(np.int64)                          # (int) truncation of a non-negative float
    p_half = ci[:-5] > 2 * ci[5:]                    # p1 > 2*s1: pulses at n, n+10, gap at n+5, n+15
    q_half = ci[5:] > 2 * ci[:-5]                    # p2 > 2*s2 seen from m=n+30: pulses at m+5, m+15
    pre = p_half[:n_end] & q_half[30:30 + n_end]
    bit = a[:-5] > a[5:]                             # a[q] > a[q+5]
    b7, b6, b5, b4, b3 = (bit[80 + 10 * k:80 + 10 * k + n_end] for k in range(5))
    m17 = pre & b7 & ~b6 & ~b5 & ~b4 & b3
    m18 = pre & b7 & ~b6 & ~b5 & b4 & ~b3 if accept_df18 else np.zeros_like(pre)
    m11 = pre & ~b7 & b6 & ~b5 & b4 & b3
    idx = np.nonzero(m17 | m18 | m11)[0]
    cls = np.where(m11[idx], 0, np.where(m17[idx], 1, 2))
    tries = np.stack([idx, cls], axis=1).astype(np.int64)
    tab = crc_table()
    out = []
    for n, k in tries:
        ln = 7 if k == 0 else 14
        pos = n + 80 + 10 * np.arange(8 * ln)
        bits = bit[pos].astype(np.uint8)
        msg = np.packbits(bits)
        r = 0
        for v in msg[:ln - 3]:
            r = ((r << 8) ^ int(tab[int(v) ^ ((r >> 16) & 0xFF)])) & 0xFFFFFF
        r ^= (int(msg[ln - 3]) << 16) | (int(msg[ln - 2]) << 8) | int(msg[ln - 1])
        if r == 0:
            pw = (int(ci[n]) + int(ci[n + 35])) // 4
            rec = np.zeros((), dtype=CAND_DTYPE)
            rec["n"] = n
            rec["pw"] = pw
            rec["len"] = ln
            rec["bytes"][:ln] = msg
            out.append(rec)
    cands = np.array(out, dtype=CAND_DTYPE) if out else np.zeros(0, dtype=CAND_DTYPE)
    return cands, tries
