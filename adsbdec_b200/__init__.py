"""B200-native Mode S demodulation path, drop-in for ``adsbdec -f`` (TLeconte/adsbdec).

The product is ``libadsbdec_b200.so`` (hand-written sm_100a CUDA kernels behind the C ABI of
``include/adsbdec_b200.h``) and the ``adsbdec_b200`` command line.  This module is the thin
ctypes host used by tests and ``bench.py``; PyTorch is only plumbing (device buffers, streams,
``torch.distributed``).  There is no CPU fallback: constructing a :class:`Decoder` without the
library or without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from dataclasses import dataclass
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libadsbdec_b200.so"
CLI_PATH = PKG_DIR / "adsbdec_b200"
CSRC = PKG_DIR / "csrc"

ABI_VERSION = 1
SINK_FN = C.CFUNCTYPE(None, C.POINTER(C.c_uint8), C.c_int, C.c_uint64, C.c_uint32, C.c_void_p)

# every symbol include/adsbdec_b200.h declares (checked by the CPU test-suite)
EXPORTS = [
    "adsbdec_b200_default_config", "adsbdec_b200_open", "adsbdec_b200_close", "adsbdec_b200_reset",
    "adsbdec_b200_set_sink",
    "adsbdec_b200_push", "adsbdec_b200_push_device", "adsbdec_b200_finish", "adsbdec_b200_decode_file",
    "adsbdec_b200_flush", "adsbdec_b200_poll", "adsbdec_b200_push_packed12", "adsbdec_b200_repair_lookup",
    "adsbdec_b200_stats", "adsbdec_b200_num_frames", "adsbdec_b200_frames", "adsbdec_b200_clear_frames",
    "adsbdec_b200_kernel_time", "adsbdec_b200_telemetry", "adsbdec_b200_fused_time", "adsbdec_b200_gate_time",
    "adsbdec_b200_format_packet", "adsbdec_b200_format_frames",
    "adsbdec_b200_envelope", "adsbdec_b200_resolver_new", "adsbdec_b200_resolver_free",
    "adsbdec_b200_resolver_feed", "adsbdec_b200_resolver_feed_frames", "adsbdec_b200_resolver_finish",
    "adsbdec_b200_resolver_advance_keys",
    "adsbdec_b200_resolver_scan_end", "adsbdec_b200_resolver_num_frames", "adsbdec_b200_resolver_frames",
    "adsbdec_b200_chunk_candidates", "adsbdec_b200_debug_fused", "adsbdec_b200_last_error",
    "adsbdec_b200_version",
]


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("accept_df18", C.c_int32),
                ("emulate_reference_tail", C.c_int32), ("collect_stats", C.c_int32), ("kernel", C.c_int32),
                ("batch_samples", C.c_uint64), ("start_sample", C.c_uint64), ("reserved", C.c_uint64 * 4)]


class Telemetry(C.Structure):
    _fields_ = [("wait_job_ms", C.c_double), ("resolve_ms", C.c_double), ("stats_issue_ms", C.c_double),
                ("sink_ms", C.c_double), ("wait_upload_ms", C.c_double), ("jobs", C.c_uint64),
                ("candidates", C.c_uint64), ("tries", C.c_uint64)]


class CFrame(C.Structure):
    _fields_ = [("ts", C.c_uint64), ("n", C.c_uint64), ("pw", C.c_uint32), ("len", C.c_int32),
                ("bytes", C.c_uint8 * 14), ("pad", C.c_uint8 * 2)]


class CCandidate(C.Structure):
    _fields_ = [("n", C.c_uint64), ("pw", C.c_uint32), ("len", C.c_uint32), ("bytes", C.c_uint8 * 16)]


CAND_DTYPE = np.dtype([("n", "<u8"), ("pw", "<u4"), ("len", "<u4"), ("bytes", "u1", (16,))])
FRAME_DTYPE = np.dtype([("ts", "<u8"), ("n", "<u8"), ("pw", "<u4"), ("len", "<i4"), ("bytes", "u1", (14,)),
                        ("pad", "u1", (2,))])


@dataclass(frozen=True)
class Frame:
    n: int
    ts: int
    pw: int
    data: bytes

    @property
    def len(self) -> int:
        return len(self.data)


class AdsbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"adsbdec_b200 error {code}: {msg}")
        self.code = code


def build(verbose: bool = False) -> None:
    """Compile the CUDA library and the CLI in-tree (nvcc, sm_100a)."""
    r = subprocess.run(["make", "-C", str(CSRC), "all"], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode:
        raise RuntimeError("building libadsbdec_b200.so failed")


_lib = None


def lib():
    """The loaded C-ABI library.  Raises if it has not been built - never falls back."""
    global _lib
    if _lib is None:
        import os
        path = Path(os.environ.get("ADSB_B200_LIB", LIB_PATH))  # kernel A/B runs (tools/try_variants.sh)
        if not path.exists():
            raise FileNotFoundError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; "
                                    "g.build()'` (there is no CPU fallback)")
        L = C.CDLL(str(path.resolve()))
        vp, u64, u32p = C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32)
        L.adsbdec_b200_default_config.argtypes = [C.POINTER(Config)]
        L.adsbdec_b200_default_config.restype = None
        L.adsbdec_b200_open.argtypes = [C.POINTER(Config), C.POINTER(vp)]
        L.adsbdec_b200_close.argtypes = [vp]
        L.adsbdec_b200_close.restype = None
        L.adsbdec_b200_reset.argtypes = [vp, u64]
        L.adsbdec_b200_set_sink.argtypes = [vp, SINK_FN, vp]
        L.adsbdec_b200_push.argtypes = [vp, vp, C.c_size_t]
        L.adsbdec_b200_push_device.argtypes = [vp, vp, C.c_size_t]
        L.adsbdec_b200_finish.argtypes = [vp]
        L.adsbdec_b200_flush.argtypes = [vp]
        L.adsbdec_b200_repair_lookup.argtypes = [C.c_uint32]
        L.adsbdec_b200_repair_lookup.restype = C.c_int
        L.adsbdec_b200_push_packed12.argtypes = [vp, vp, C.c_size_t]
        L.adsbdec_b200_poll.argtypes = [vp]
        L.adsbdec_b200_decode_file.argtypes = [vp, C.c_char_p]
        L.adsbdec_b200_stats.argtypes = [vp, u32p, u32p]
        L.adsbdec_b200_num_frames.argtypes = [vp]
        L.adsbdec_b200_num_frames.restype = C.c_size_t
        L.adsbdec_b200_frames.argtypes = [vp]
        L.adsbdec_b200_frames.restype = vp
        L.adsbdec_b200_clear_frames.argtypes = [vp]
        L.adsbdec_b200_clear_frames.restype = None
        L.adsbdec_b200_kernel_time.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(u64), C.c_int]
        L.adsbdec_b200_telemetry.argtypes = [vp, C.POINTER(Telemetry), C.c_int]
        L.adsbdec_b200_fused_time.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(u64), C.POINTER(u64), C.c_int]
        L.adsbdec_b200_format_frames.argtypes = [C.c_int, vp, C.c_size_t, vp, C.c_size_t]
        L.adsbdec_b200_format_frames.restype = C.c_int64
        L.adsbdec_b200_gate_time.argtypes = [vp, C.POINTER(C.c_double), C.c_int]
        L.adsbdec_b200_format_packet.argtypes = [C.c_int, C.c_char_p, C.c_int, u64, C.c_uint32, C.c_char_p]
        L.adsbdec_b200_envelope.argtypes = [C.c_int, vp, C.c_size_t, C.c_int, u64, vp, C.c_size_t]
        L.adsbdec_b200_resolver_new.argtypes = [C.c_int]
        L.adsbdec_b200_resolver_new.restype = vp
        L.adsbdec_b200_resolver_free.argtypes = [vp]
        L.adsbdec_b200_resolver_free.restype = None
        L.adsbdec_b200_resolver_feed.argtypes = [vp, vp, C.c_size_t, u64, u64, SINK_FN, vp]
        L.adsbdec_b200_resolver_feed_frames.argtypes = [vp, vp, C.c_size_t, u64, u64, u32p, SINK_FN, vp]
        L.adsbdec_b200_resolver_finish.argtypes = [vp, SINK_FN, vp]
        L.adsbdec_b200_resolver_advance_keys.argtypes = [vp, vp, C.c_size_t, C.c_size_t, u64, u64, C.c_int]
        L.adsbdec_b200_resolver_advance_keys.restype = C.c_int64
        L.adsbdec_b200_resolver_scan_end.argtypes = [vp]
        L.adsbdec_b200_resolver_scan_end.restype = u64
        L.adsbdec_b200_chunk_candidates.argtypes = [vp, vp, C.c_size_t, u64, u64, u64, vp, C.c_size_t, u32p]
        L.adsbdec_b200_chunk_candidates.restype = C.c_int64
        L.adsbdec_b200_resolver_num_frames.argtypes = [vp]
        L.adsbdec_b200_resolver_num_frames.restype = C.c_size_t
        L.adsbdec_b200_resolver_frames.argtypes = [vp]
        L.adsbdec_b200_resolver_frames.restype = vp
        L.adsbdec_b200_debug_fused.argtypes = [vp, vp, C.c_size_t, u64, vp, vp, vp, vp, C.c_size_t,
                                               C.POINTER(C.c_int64)]
        L.adsbdec_b200_last_error.restype = C.c_char_p
        L.adsbdec_b200_version.restype = C.c_char_p
        _lib = L
    return _lib


def _check(rc: int) -> int:
    if rc < 0:
        raise AdsbError(rc, lib().adsbdec_b200_last_error().decode(errors="replace"))
    return rc


def format_packet(fmt: int, frame: bytes, ts: int, pw: int) -> bytes:
    """formatpkt() (output.c:204-262): fmt 0 avr, 1 avrmlat, 2 beast."""
    buf = C.create_string_buffer(256)
    n = _check(lib().adsbdec_b200_format_packet(fmt, frame, len(frame), ts, pw, buf))
    return buf.raw[:n]


def format_frames(fmt: int, frames: np.ndarray) -> bytes:
    """formatpkt() over a FRAME_DTYPE array, packets back to back (what runOutput() writes)."""
    a = np.ascontiguousarray(frames, dtype=FRAME_DTYPE)
    out = np.empty(48 * max(1, a.size), dtype=np.uint8)
    n = _check(int(lib().adsbdec_b200_format_frames(fmt, a.ctypes.data, a.size, out.ctypes.data, out.size)))
    return out[:n].tobytes()


def _frames_from_array(arr: np.ndarray) -> list[Frame]:
    return [Frame(int(r["n"]), int(r["ts"]), int(r["pw"]), bytes(r["bytes"][:int(r["len"])])) for r in arr]


class Decoder:
    """One stream through the GPU path; mirrors the reference's per-stream state
    (air.c:33-34, :49-50; demod.c:86; valid.c:30-31).

    >>> d = Decoder(accept_df18=False)
    >>> d.push(samples_u16); d.finish(); d.frames
    """

    def __init__(self, *, device: int = 0, accept_df18: bool = False, emulate_reference_tail: bool = True,
                 collect_stats: bool = True, kernel: int = 0, batch_samples: int = 0, start_sample: int = 0,
                 sink=None, initial_caps: tuple[int, int] = (0, 0), reader_threads: int = 0, repair: int = 0,
                 address_parity: bool = False, iq: bool = False):
        L = lib()
        cfg = Config()
        L.adsbdec_b200_default_config(C.byref(cfg))
        cfg.device = device
        cfg.accept_df18 = int(accept_df18)
        cfg.emulate_reference_tail = int(emulate_reference_tail)
        cfg.collect_stats = int(collect_stats)
        cfg.kernel = kernel
        cfg.batch_samples = batch_samples
        cfg.start_sample = start_sample
        cfg.reserved[0], cfg.reserved[1] = initial_caps
        cfg.reserved[2] = reader_threads  # decode_file: 0 = up to 16, 1 = the calling thread only
        # EXTENSIONS (not in the reference): repair 1 = fix one bit, 2 = two bits; 0x100 = accept DF4/5/20/21
        # whose address/parity field names an aircraft announced earlier in the stream
        # 0x200 = the samples are signed I/Q pairs at 10 MS/s (complex baseband) instead of real samples at 20 MS/s
        cfg.reserved[3] = repair | (0x100 if address_parity else 0) | (0x200 if iq else 0)
        self._h = C.c_void_p()
        self._cb = None
        _check(L.adsbdec_b200_open(C.byref(cfg), C.byref(self._h)))
        if sink is not None:
            self.set_sink(sink)

    def set_sink(self, fn) -> None:
        """fn(frame: bytes, ts: int, pw: int), called like netout() (output.c:159)."""
        def tramp(p, n, ts, pw, _user):
            fn(bytes(p[:n]), int(ts), int(pw))
        self._cb = SINK_FN(tramp)
        _check(lib().adsbdec_b200_set_sink(self._h, self._cb, None))

    def reset(self, start_sample: int = 0) -> None:
        """Back to an empty stream, buffers kept."""
        _check(lib().adsbdec_b200_reset(self._h, start_sample))

    def push(self, samples: np.ndarray) -> None:
        a = np.ascontiguousarray(samples, dtype="<u2")
        _check(lib().adsbdec_b200_push(self._h, a.ctypes.data, a.size))

    def push_packed12(self, packed: np.ndarray, n_samples: int | None = None) -> None:
        """Samples packed 8 to 12 bytes (see pack12); a quarter less PCIe traffic than push()."""
        a = np.ascontiguousarray(packed).view(np.uint8)
        n = a.size // 3 * 2 if n_samples is None else n_samples
        _check(lib().adsbdec_b200_push_packed12(self._h, a.ctypes.data, n))

    def push_device(self, dev_ptr: int, n_samples: int) -> None:
        _check(lib().adsbdec_b200_push_device(self._h, C.c_void_p(dev_ptr), n_samples))

    def finish(self) -> None:
        _check(lib().adsbdec_b200_finish(self._h))

    def flush(self) -> int:
        """Submit the staged samples as a scan job now (live sources); returns frames delivered."""
        rc = lib().adsbdec_b200_flush(self._h)
        _check(rc)
        return rc

    def poll(self) -> int:
        """Deliver the frames of finished scan jobs without blocking; returns how many."""
        rc = lib().adsbdec_b200_poll(self._h)
        _check(rc)
        return rc

    def decode_file(self, path) -> int:
        """fileInput() (air.c:219-246).  Returns the C return code; a missing file is
        ADSB_B200_EIO (-6) and yields no frames, as in the reference."""
        rc = lib().adsbdec_b200_decode_file(self._h, str(path).encode())
        if rc < 0 and rc != -6:
            _check(rc)
        return rc

    def chunk_candidates(self, dev_ptr: int, n_samples: int, start_sample: int, own_begin: int, own_end: int,
                         cap: int = 1 << 20):
        """Every CRC-valid candidate starting in [own_begin, own_end) (chunk-local offsets)."""
        out = np.zeros(cap, dtype=CAND_DTYPE)
        tries = (C.c_uint32 * 3)()
        n = lib().adsbdec_b200_chunk_candidates(self._h, C.c_void_p(dev_ptr), n_samples, start_sample,
                                                own_begin, own_end, out.ctypes.data, cap, tries)
        _check(int(n))
        return out[:n], tuple(tries)

    def debug_fused(self, samples: np.ndarray, count: int, start_sample: int = 0):
        """(j_begin, amp, p_bits, q_bits, b_bits) as the fused kernel computed them (parity tap)."""
        a = np.ascontiguousarray(samples, dtype="<u2")
        count -= count % 32
        amp = np.zeros(count, dtype=np.float32)
        pw = np.zeros(count // 32, dtype=np.uint32)
        qw = np.zeros(count // 32, dtype=np.uint32)
        bw = np.zeros(count // 32, dtype=np.uint32)
        jb = C.c_int64()
        _check(lib().adsbdec_b200_debug_fused(self._h, a.ctypes.data, a.size, start_sample, amp.ctypes.data,
                                              pw.ctypes.data, qw.ctypes.data, bw.ctypes.data, count,
                                              C.byref(jb)))
        unpack = lambda w: np.unpackbits(w.astype(">u4").view(np.uint8)).astype(bool)
        return jb.value, amp, unpack(pw), unpack(qw), unpack(bw)

    @property
    def frames_array(self) -> np.ndarray:
        L = lib()
        n = L.adsbdec_b200_num_frames(self._h)
        if n == 0:
            return np.zeros(0, dtype=FRAME_DTYPE)
        p = L.adsbdec_b200_frames(self._h)
        buf = (C.c_uint8 * (n * FRAME_DTYPE.itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=FRAME_DTYPE).copy()

    @property
    def frames_view(self) -> np.ndarray:
        """The frame store itself (adsbdec_b200_frames()), not a copy: page-locked host memory the GPU has written
        the accepted frames into.  Valid until the next push / finish / reset / clear_frames / close."""
        L = lib()
        n = L.adsbdec_b200_num_frames(self._h)
        if n == 0:
            return np.zeros(0, dtype=FRAME_DTYPE)
        buf = (C.c_uint8 * (n * FRAME_DTYPE.itemsize)).from_address(L.adsbdec_b200_frames(self._h))
        return np.frombuffer(buf, dtype=FRAME_DTYPE)

    @property
    def num_frames(self) -> int:
        return int(lib().adsbdec_b200_num_frames(self._h))

    @property
    def frames(self) -> list[Frame]:
        return _frames_from_array(self.frames_array)

    def stats(self):
        t = (C.c_uint32 * 3)()
        k = (C.c_uint32 * 3)()
        _check(lib().adsbdec_b200_stats(self._h, t, k))
        return tuple(t), tuple(k)

    def kernel_time(self, reset: bool = False):
        ms = C.c_double()
        n = C.c_uint64()
        _check(lib().adsbdec_b200_kernel_time(self._h, C.byref(ms), C.byref(n), int(reset)))
        return ms.value, n.value

    def host_time(self, reset: bool = False) -> dict:
        """Host-side telemetry: wall ms per phase of the caller's thread plus what the kernels reported
        (jobs, CRC-valid candidates, gate-passing offsets)."""
        t = Telemetry()
        _check(lib().adsbdec_b200_telemetry(self._h, C.byref(t), int(reset)))
        return {name: getattr(t, name) for name, _ in Telemetry._fields_}

    def fused_time(self, reset: bool = False):
        ms = C.c_double()
        n = C.c_uint64()
        s = C.c_uint64()
        _check(lib().adsbdec_b200_fused_time(self._h, C.byref(ms), C.byref(n), C.byref(s), int(reset)))
        return ms.value, n.value, s.value

    def gate_time(self, reset: bool = False) -> float:
        ms = C.c_double()
        _check(lib().adsbdec_b200_gate_time(self._h, C.byref(ms), int(reset)))
        return ms.value

    def text(self, fmt: int = 0) -> bytes:
        return b"".join(format_packet(fmt, f.data, f.ts, f.pw) for f in self.frames)

    def close(self) -> None:
        if self._h:
            lib().adsbdec_b200_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def repair_lookup(syndrome: int) -> int:
    """Bits the repair extension flips for this CRC syndrome: (first << 8) | second, 0xFF = none; -1 if unknown."""
    return int(lib().adsbdec_b200_repair_lookup(syndrome))


def pack12(samples: np.ndarray) -> np.ndarray:
    """12-bit samples (u16, < 4096, a multiple of 8 of them) packed the way the airspy does with packing on
    (libairspy unpack_samples): the 96 bits of three little-endian 32-bit words, most significant first, are
    eight samples.  Returns the words as a uint32 array."""
    s = np.ascontiguousarray(samples, dtype="<u2").astype(np.uint32).reshape(-1, 8)
    if (s > 0xFFF).any():
        raise ValueError("samples do not fit 12 bits")
    w = np.empty((s.shape[0], 3), dtype="<u4")
    w[:, 0] = (s[:, 0] << 20) | (s[:, 1] << 8) | (s[:, 2] >> 4)
    w[:, 1] = ((s[:, 2] & 0xF) << 28) | (s[:, 3] << 16) | (s[:, 4] << 4) | (s[:, 5] >> 8)
    w[:, 2] = ((s[:, 5] & 0xFF) << 24) | (s[:, 6] << 12) | s[:, 7]
    return w.reshape(-1)


def unpack12(words: np.ndarray) -> np.ndarray:
    """Inverse of pack12 on the host (tests)."""
    w = np.ascontiguousarray(words, dtype="<u4").reshape(-1, 3)
    s = np.empty((w.shape[0], 8), dtype="<u2")
    s[:, 0] = w[:, 0] >> 20
    s[:, 1] = (w[:, 0] >> 8) & 0xFFF
    s[:, 2] = ((w[:, 0] & 0xFF) << 4) | (w[:, 1] >> 28)
    s[:, 3] = (w[:, 1] >> 16) & 0xFFF
    s[:, 4] = (w[:, 1] >> 4) & 0xFFF
    s[:, 5] = ((w[:, 1] & 0xF) << 8) | (w[:, 2] >> 24)
    s[:, 6] = (w[:, 2] >> 12) & 0xFFF
    s[:, 7] = w[:, 2] & 0xFFF
    return s.reshape(-1)


def decode(samples: np.ndarray, **kw) -> "Decoder":
    """Whole capture through push()/finish(); returns the finished decoder."""
    d = Decoder(**kw)
    d.push(samples)
    d.finish()
    return d


def envelope(samples: np.ndarray, count: int | None = None, *, device: int = 0, start_sample: int = 0) -> np.ndarray:
    """Envelope samples (air.c:59-92) computed on the GPU by the generic kernel (parity tap)."""
    a = np.ascontiguousarray(samples, dtype="<u2")
    count = a.size // 2 if count is None else count
    out = np.zeros(count, dtype=np.float32)
    _check(lib().adsbdec_b200_envelope(device, a.ctypes.data, a.size, 0, start_sample, out.ctypes.data, count))
    return out


class Resolver:
    """Host-only sequential rules over a sorted candidate list."""

    def __init__(self, emulate_reference_tail: bool = True, keep: bool = False):
        """keep=True: no per-frame callback; the frames pile up in the C resolver (frames_array)."""
        self._r = lib().adsbdec_b200_resolver_new(int(emulate_reference_tail))
        self.out: list[Frame] = []

        def tramp(p, n, ts, pw, _u):
            self.out.append(Frame(-1, int(ts), int(pw), bytes(p[:n])))
        self._cb = SINK_FN() if keep else SINK_FN(tramp)

    @property
    def frames_array(self) -> np.ndarray:
        L = lib()
        n = L.adsbdec_b200_resolver_num_frames(self._r)
        if n == 0:
            return np.zeros(0, dtype=FRAME_DTYPE)
        buf = (C.c_uint8 * (n * FRAME_DTYPE.itemsize)).from_address(L.adsbdec_b200_resolver_frames(self._r))
        return np.frombuffer(buf, dtype=FRAME_DTYPE).copy()

    def feed(self, cands: np.ndarray, scanned_upto: int, total_out: int) -> None:
        c = np.ascontiguousarray(cands, dtype=CAND_DTYPE)
        _check(lib().adsbdec_b200_resolver_feed(self._r, c.ctypes.data, c.size, scanned_upto, total_out,
                                                self._cb, None))

    def feed_frames(self, frames: np.ndarray, scanned_upto: int, total_out: int):
        """Frames that already carry the greedy decision and their ts (FRAME_DTYPE): only the deqframe call
        schedule is applied.  Returns the Ok counts (DF11, DF17, DF18) of what was accepted by this call."""
        f = np.ascontiguousarray(frames, dtype=FRAME_DTYPE)
        ok = (C.c_uint32 * 3)()
        _check(lib().adsbdec_b200_resolver_feed_frames(self._r, f.ctypes.data, f.size, scanned_upto, total_out,
                                                       ok, self._cb, None))
        return tuple(ok)

    def advance_keys(self, keys: np.ndarray, n_vis: int, scanned_upto: int, total_out: int, finish: bool = False) -> int:
        """The in-place schedule of the stream decoder: keys[i] = start << 1 | long for ALL stamped frames so far,
        the first n_vis of them reached before.  Returns the new n_vis."""
        k = np.ascontiguousarray(keys, dtype=np.uint64)
        return _check(int(lib().adsbdec_b200_resolver_advance_keys(self._r, k.ctypes.data, n_vis, k.size, scanned_upto,
                                                                  total_out, int(finish))))

    def finish(self) -> None:
        _check(lib().adsbdec_b200_resolver_finish(self._r, self._cb, None))

    @property
    def scan_end(self) -> int:
        return int(lib().adsbdec_b200_resolver_scan_end(self._r))

    def __del__(self):
        try:
            lib().adsbdec_b200_resolver_free(self._r)
        except Exception:
            pass
