"""Time-chunking of a long capture across decoders / GPUs - host logic only.

The hot path shards by TIME: every chunk owns a range of preamble start offsets, carries a halo of
one maximum-length frame plus preamble (1200 envelope samples, adsbdec.h:3 DECOFFSET) behind it
and a few samples of FIR history in front, and is scanned independently (no data-path
collective).  The per-chunk candidate lists are concatenated in offset order and the sequential
rules of demod.c:84-144 / air.c:94-99 (greedy skip, iteration-count timestamp, unscanned tail)
are applied ONCE over the merged list by the host resolver - so the merged output is what a
single decoder would have produced for the whole capture, duplicates cannot arise (ownership is
by start offset), and nothing depends on where the cuts fall.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

LOOKAHEAD = 1200          # envelope samples a candidate reads past its start (adsbdec.h:3)
LEAD_OUT = 16             # envelope samples of real FIR history in front of the owned range


@dataclass(frozen=True)
class Chunk:
    index: int
    sample_begin: int     # first capture sample the chunk needs (multiple of 8: 16-byte aligned)
    sample_end: int       # one past the last sample it needs
    own_begin: int        # owned offsets, ABSOLUTE envelope-sample indices
    own_end: int

    @property
    def out_base(self) -> int:
        """Absolute offset of the chunk's local offset 0."""
        return self.sample_begin // 2

    @property
    def local_own(self) -> tuple[int, int]:
        return self.own_begin - self.out_base, self.own_end - self.out_base


def plan_chunks(n_samples: int, n_chunks: int) -> list[Chunk]:
    """Cut [0, n_samples) into n_chunks time chunks with halo.  Offsets >= total_out-1200 have no
    complete look-ahead and are owned by nobody (the reference never scans them either)."""
    total_out = n_samples // 2
    scan_end = max(0, total_out - LOOKAHEAD)
    n_chunks = max(1, min(n_chunks, max(1, scan_end // 4096)))
    bounds = [(scan_end * i // n_chunks) // 4 * 4 for i in range(n_chunks)] + [scan_end]
    chunks = []
    for i in range(n_chunks):
        a, b = bounds[i], bounds[i + 1]
        s0 = max(0, 2 * (a - LEAD_OUT)) // 8 * 8
        s1 = min(n_samples, 2 * (b + LOOKAHEAD))
        chunks.append(Chunk(i, s0, s1, a, b))
    return chunks


def assign(chunks: list[Chunk], world_size: int, rank: int) -> list[Chunk]:
    """Round-robin chunk -> rank map (each rank scans its chunks on its own GPU)."""
    return [c for c in chunks if c.index % world_size == rank]


def merge_candidates(parts: list[tuple[Chunk, np.ndarray]]) -> np.ndarray:
    """Concatenate per-chunk candidate arrays (chunk-LOCAL offsets) into one array with absolute
    offsets, sorted by offset.  Ownership ranges are disjoint, so no de-duplication is needed; it
    is asserted."""
    from . import CAND_DTYPE
    out = []
    for ch, c in sorted(parts, key=lambda p: p[0].own_begin):
        c = np.array(c, dtype=CAND_DTYPE, copy=True)
        c["n"] += ch.out_base
        if c.size:
            assert c["n"].min() >= ch.own_begin and c["n"].max() < ch.own_end, "candidate outside its chunk"
        out.append(c)
    merged = np.concatenate(out) if out else np.zeros(0, dtype=CAND_DTYPE)
    # every chunk's list comes sorted and the ownership ranges are disjoint and taken in order: the
    # concatenation IS the merge (a sort of millions of records would cost more than the scans)
    if merged.size > 1 and not np.all(merged["n"][1:] >= merged["n"][:-1]):
        merged = merged[np.argsort(merged["n"], kind="stable")]
    return merged


def resolve_merged(merged: np.ndarray, n_samples: int, *, emulate_reference_tail: bool = True, as_array: bool = False):
    """Sequential rules over the merged list; returns the accepted frames (ts, pw, bytes) - as a list of
    Frame, or (as_array) as a FRAME_DTYPE array without a Python call per frame."""
    from . import Resolver
    total_out = n_samples // 2
    r = Resolver(emulate_reference_tail, keep=as_array)
    r.feed(merged, max(0, total_out - LOOKAHEAD), total_out)
    r.finish()
    return r.frames_array if as_array else r.out
