"""ctypes front for oracle/liboracle.so and a runner for oracle/_ref binaries.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(adsbdec_b200/) must never import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import socket
import subprocess
import threading
from dataclasses import dataclass
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle.so"
REF_DIR = HERE / "_ref"


class _Frame(C.Structure):
    _fields_ = [("ts", C.c_uint64), ("n", C.c_uint64), ("pw", C.c_uint32), ("len", C.c_int32),
                ("bytes", C.c_uint8 * 14), ("pad", C.c_uint8 * 2)]


@dataclass(frozen=True)
class Frame:
    n: int
    ts: int
    pw: int
    data: bytes

    @property
    def len(self) -> int:
        return len(self.data)


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    if force or not LIB.exists() or LIB.stat().st_mtime < (HERE / "oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "port"], check=True, capture_output=True)
    if Path("/root/reference/air.c").exists():
        subprocess.run(["make", "-C", str(HERE), "ref"], check=True, capture_output=True)
        if (HERE.parent / "adsbdec_b200" / "libadsbdec_b200.so").exists():
            # the reference with INTEGRATION.md's patch applied, linked against the product library
            subprocess.run(["make", "-C", str(HERE), "patched"], check=True, capture_output=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB))
        L.oracle_new.restype = C.c_void_p
        L.oracle_new.argtypes = [C.c_int, C.c_int]
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_set_origin.argtypes = [C.c_void_p, C.c_uint64]
        L.oracle_set_repair.argtypes = [C.c_void_p, C.c_int]
        L.oracle_set_address_parity.argtypes = [C.c_void_p, C.c_int]
        L.oracle_set_iq.argtypes = [C.c_void_p, C.c_int]
        L.oracle_address_prepass.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.oracle_push.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_decode_file.argtypes = [C.c_void_p, C.c_char_p]
        L.oracle_decode_file.restype = C.c_int
        L.oracle_decode_memory.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.oracle_num_frames.argtypes = [C.c_void_p]
        L.oracle_num_frames.restype = C.c_size_t
        L.oracle_frames.argtypes = [C.c_void_p]
        L.oracle_frames.restype = C.POINTER(_Frame)
        L.oracle_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.oracle_record_amp.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.oracle_amp_count.argtypes = [C.c_void_p]
        L.oracle_amp_count.restype = C.c_size_t
        L.oracle_format.argtypes = [C.POINTER(_Frame), C.c_int, C.c_char_p]
        L.oracle_format.restype = C.c_int
        L.oracle_crc_table.restype = C.POINTER(C.c_uint32)
        _lib = L
    return _lib


@dataclass
class Result:
    frames: list[Frame]
    tried: tuple[int, int, int]
    ok: tuple[int, int, int]
    amp: np.ndarray | None = None
    fixed: list[int] | None = None     # bits repaired per frame (extension; all 0 by default)

    def text(self, fmt: int = 0) -> bytes:
        return b"".join(format_frame(f, fmt) for f in self.frames)


def format_frame(f: Frame, fmt: int) -> bytes:
    L = lib()
    cf = _Frame(ts=f.ts, n=f.n, pw=f.pw, len=len(f.data))
    for i, b in enumerate(f.data):
        cf.bytes[i] = b
    buf = C.create_string_buffer(256)
    n = L.oracle_format(C.byref(cf), fmt, buf)
    return buf.raw[:n]


def decode(samples: np.ndarray, *, accept_df18: bool = False, fma: bool = True,
           want_amp: bool = False, origin: int = 0, repair: int = 0, address_parity: bool = False,
           iq: bool = False) -> Result:
    """Run the sequential restatement over a whole capture (file semantics)."""
    L = lib()
    samples = np.ascontiguousarray(samples, dtype="<u2")
    ctx = L.oracle_new(int(accept_df18), int(fma))
    amp = None
    try:
        if origin:
            L.oracle_set_origin(ctx, origin)
        if repair:
            L.oracle_set_repair(ctx, repair)   # extension, see oracle.h
        if iq:                                 # extension, see oracle.h
            L.oracle_set_iq(ctx, 1)
        if address_parity:                     # extension, see oracle.h
            L.oracle_set_address_parity(ctx, 1)
            L.oracle_address_prepass(ctx, samples.ctypes.data, samples.nbytes)
        if want_amp:
            amp = np.zeros(samples.size // 2 + 8, dtype=np.float32)
            L.oracle_record_amp(ctx, amp.ctypes.data, amp.size)
        L.oracle_decode_memory(ctx, samples.ctypes.data, samples.nbytes)
        n = L.oracle_num_frames(ctx)
        fr = L.oracle_frames(ctx)
        frames = [Frame(int(fr[i].n), int(fr[i].ts), int(fr[i].pw), bytes(fr[i].bytes[:fr[i].len]))
                  for i in range(n)]
        t = (C.c_uint32 * 3)()
        k = (C.c_uint32 * 3)()
        L.oracle_stats(ctx, t, k)
        if want_amp:
            amp = amp[:L.oracle_amp_count(ctx)]
        res = Result(frames, tuple(t), tuple(k), amp)
        res.fixed = [int(fr[i].pad[0]) for i in range(n)]
        return res
    finally:
        L.oracle_free(ctx)


def decode_file(path, *, accept_df18: bool = False, fma: bool = True) -> Result:
    """fileInput() semantics on a file as it is on disk (2 MiB reads through one reused buffer, ragged
    and odd-length tails included)."""
    L = lib()
    ctx = L.oracle_new(int(accept_df18), int(fma))
    try:
        if L.oracle_decode_file(ctx, str(path).encode()) != 0:
            raise FileNotFoundError(path)
        n = L.oracle_num_frames(ctx)
        fr = L.oracle_frames(ctx)
        frames = [Frame(int(fr[i].n), int(fr[i].ts), int(fr[i].pw), bytes(fr[i].bytes[:fr[i].len]))
                  for i in range(n)]
        t = (C.c_uint32 * 3)()
        k = (C.c_uint32 * 3)()
        L.oracle_stats(ctx, t, k)
        return Result(frames, tuple(t), tuple(k), None)
    finally:
        L.oracle_free(ctx)


def crc_table() -> np.ndarray:
    return np.ctypeslib.as_array(lib().oracle_crc_table(), shape=(256,)).copy()


# ---- the real reference binary (oracle/_ref) --------------------------------

def ref_binary(flavour: str | None = None) -> Path | None:
    """A runnable oracle/_ref/adsbdec_* binary, or None.

    ``native`` was built with the reference's own -march=native on the build
    host and may not run elsewhere; ``v3`` has the same FMA-contracted arithmetic
    and runs on any AVX2+FMA machine.
    """
    order = [flavour] if flavour else ["native", "v3"]
    for fl in order:
        p = REF_DIR / f"adsbdec_{fl}"
        if not p.exists():
            continue
        try:
            r = subprocess.run([str(p), "-f", "/nonexistent/x"], capture_output=True, timeout=20)
            if r.returncode == 0:
                return p
        except (OSError, subprocess.SubprocessError):
            pass
    return None


@dataclass
class RefRun:
    stdout: bytes
    stderr: bytes
    returncode: int

    def stats(self):
        """(try, ok) triplets parsed from print_stats (valid.c:84-101)."""
        t = k = None
        for line in self.stderr.decode(errors="replace").splitlines():
            if line.startswith("Try :"):
                t = tuple(int(x) for x in line.split(":")[1].split())
            elif line.startswith("Ok :"):
                k = tuple(int(x) for x in line.split(":")[1].split())
        return t, k


def run_ref(path, flags=(), binary: Path | None = None, timeout: float = 600) -> RefRun:
    """``adsbdec <flags> -f path`` with stdout captured (avr / avrmlat)."""
    binary = binary or ref_binary()
    if binary is None:
        raise FileNotFoundError("oracle/_ref not built")
    r = subprocess.run([str(binary), *flags, "-f", str(path)], capture_output=True, timeout=timeout)
    return RefRun(r.stdout, r.stderr, r.returncode)


def run_ref_tcp(path, flags=("-b",), binary: Path | None = None, timeout: float = 600) -> RefRun:
    """Run with ``-s 127.0.0.1:port`` into a local sink: the only way the
    reference writes intact beast packets (stdout uses strlen)."""
    binary = binary or ref_binary()
    if binary is None:
        raise FileNotFoundError("oracle/_ref not built")
    srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
    srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
    srv.bind(("127.0.0.1", 0))
    srv.listen(1)
    port = srv.getsockname()[1]
    chunks: list[bytes] = []

    def sink():
        srv.settimeout(timeout)
        try:
            conn, _ = srv.accept()
        except OSError:
            return
        with conn:
            while True:
                d = conn.recv(1 << 16)
                if not d:
                    break
                chunks.append(d)

    th = threading.Thread(target=sink, daemon=True)
    th.start()
    r = subprocess.run([str(binary), *flags, "-s", f"127.0.0.1:{port}", "-f", str(path)],
                       capture_output=True, timeout=timeout)
    th.join(timeout=10)
    srv.close()
    return RefRun(b"".join(chunks), r.stderr, r.returncode)


def run_amp_probe(path, out_path, origin: int = 0) -> np.ndarray:
    """Envelope stream of the UNMODIFIED air.c (optionally with fidx preset to origin)."""
    p = REF_DIR / "amp_probe"
    subprocess.run([str(p), str(path), str(out_path), str(origin)], check=True, timeout=600)
    return np.fromfile(out_path, dtype=np.float32)


def host_cores() -> int:
    return len(os.sched_getaffinity(0))
