"""Generates tests/golden/*.npz with the UNMODIFIED reference binaries of oracle/_ref.

Run in the build container (needs /root/reference to have been compiled by `make -C oracle ref`):

    python tests/golden/make_golden.py

Each fixture holds a small synthetic capture (uint16 samples, exactly the bytes the reference
read from disk) and what the reference printed for it:
    avr        adsbdec -f F                      (stdout)
    avrmlat    adsbdec -m -f F                   (stdout)
    avrmlat_a  adsbdec -a -m -f F                (stdout)
    beast_a    adsbdec -a -b -s 127.0.0.1:P -f F (TCP sink; stdout beast is truncated)
    stats / stats_a   Try / Ok lines of print_stats (stderr)
    amp_sha256 sha256 of the float32 envelope stream of the unmodified air.c (oracle/_ref/amp_probe),
    amp_head   its first 4096 samples
All captures end in mid-scale padding so that the reference's EOF race cannot drop
queued frames, except `ragged_tail`, whose length is deliberately 2 bytes short of a multiple of 8.
"""
import hashlib
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from adsbdec_b200.modulator import synth_capture  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

HERE = Path(__file__).resolve().parent
PAD = 1 << 15

RECIPES = {
    # config 1 shape: DF17 at 20 dB
    "df17_20db": dict(seconds=0.004, n_frames=14, snr_db=20.0, seed=101, df_mix=((17, 1.0),), poisson=False),
    # config 3 shape: DF11/17/18/20/21 with 1- and 2-bit errors; -a adds DF18
    "mix_errors": dict(seconds=0.005, n_frames=24, snr_db=22.0, seed=102,
                       df_mix=((17, 0.4), (11, 0.25), (18, 0.15), (20, 0.1), (21, 0.1)), bit_errors=1,
                       overlay_for={20: True, 21: True}, poisson=False),
    # config 2 shape: overlapping / garbled frames
    "dense": dict(seconds=0.004, n_frames=60, snr_db=18.0, seed=103,
                  df_mix=((17, 0.6), (11, 0.3), (18, 0.1)), poisson=True),
    # config 4 shape: low SNR
    "snr8": dict(seconds=0.005, n_frames=20, snr_db=8.0, seed=104, df_mix=((17, 0.7), (11, 0.3)), poisson=False),
}


def complete(run, count_frames):
    """The reference can lose frames that are still queued when the input thread hits EOF
    (output.c:287 tests do_exit before draining).  The frames themselves are
    deterministic, only the loss is a race: re-run until the output holds as many frames as the
    Ok counters say were accepted."""
    for _ in range(200):
        r = run()
        t, k = r.stats()
        if count_frames(r.stdout) == sum(k):
            return r
    raise RuntimeError("reference kept dropping queued frames at EOF")


def beast_count(b: bytes) -> int:
    n = i = 0
    while i < len(b):
        assert b[i] == 0x1A and b[i + 1] in (0x32, 0x33), "beast framing"
        ln = 7 if b[i + 1] == 0x32 else 14
        i += 2
        left = 6 + 1 + ln                 # timestamp, level, payload; 0x1a doubled except the level byte
        k = 0
        while k < left:
            if b[i] == 0x1A and k != 6:
                i += 1
            i += 1
            k += 1
        n += 1
    return n


def run_all(path, samples_for_amp_len):
    binary = po.ref_binary()
    out = {}
    lines = lambda b: b.count(b"\n")
    r = complete(lambda: po.run_ref(path, binary=binary), lines)
    out["avr"] = r.stdout
    out["stats"] = np.array(r.stats(), dtype=np.int64)
    out["avrmlat"] = complete(lambda: po.run_ref(path, flags=("-m",), binary=binary), lines).stdout
    ra = complete(lambda: po.run_ref(path, flags=("-a", "-m"), binary=binary), lines)
    out["avrmlat_a"] = ra.stdout
    out["stats_a"] = np.array(ra.stats(), dtype=np.int64)
    out["beast_a"] = complete(lambda: po.run_ref_tcp(path, flags=("-a", "-b"), binary=binary), beast_count).stdout
    with tempfile.NamedTemporaryFile(suffix=".f32") as t:
        amp = po.run_amp_probe(path, t.name)
    out["amp_sha256"] = hashlib.sha256(amp.tobytes()).hexdigest()
    out["amp_len"] = amp.size
    out["amp_head"] = amp[:4096].copy()
    return out


def main():
    po.build()
    if po.ref_binary() is None:
        raise SystemExit("oracle/_ref is not built (needs /root/reference)")
    for name, kw in RECIPES.items():
        cap = synth_capture(pad=PAD, **kw)
        samples = cap.samples.astype("<u2")
        with tempfile.NamedTemporaryFile(suffix=".bin") as f:
            samples.tofile(f.name)
            out = run_all(f.name, samples.size)
        np.savez_compressed(HERE / f"{name}.npz", samples=samples,
                            **{k: (np.frombuffer(v, dtype=np.uint8) if isinstance(v, bytes) else v)
                               for k, v in out.items()})
        print(name, samples.size, "samples;", out["avrmlat_a"].count(b"\n"), "frames with -a;",
              "stats", out["stats_a"].tolist())
    # ragged tail: byte length = 2 (mod 8) -> decodeiq over-reads 3 stale entries (air.c:59-82)
    cap = synth_capture(pad=PAD, **RECIPES["df17_20db"])
    samples = cap.samples.astype("<u2")[:-3]
    with tempfile.NamedTemporaryFile(suffix=".bin") as f:
        samples.tofile(f.name)
        out = run_all(f.name, samples.size)
    np.savez_compressed(HERE / "ragged_tail.npz", samples=samples,
                        **{k: (np.frombuffer(v, dtype=np.uint8) if isinstance(v, bytes) else v)
                           for k, v in out.items()})
    print("ragged_tail", samples.size, out["avrmlat_a"].count(b"\n"))
    # known-answer: the widely published DF17 frame has syndrome 0 under crc.h's table
    print("KAT 8D4840D6202CC371C32CE0576098")


if __name__ == "__main__":
    main()
