"""CPU: the oracle (oracle/oracle.c) against the golden vectors produced by the UNMODIFIED
reference binaries (tests/golden/make_golden.py), and - when oracle/_ref is present and runs on
this host - against the reference itself on fresh captures."""
import hashlib
from pathlib import Path

import numpy as np
import pytest

from helpers import DF_MIX_ALL, capture_basic

GOLDEN = sorted((Path(__file__).parent / "golden").glob("*.npz"))


def load(p):
    z = np.load(p)
    d = {k: z[k] for k in z.files}
    for k in ("avr", "avrmlat", "avrmlat_a", "beast_a"):
        d[k] = d[k].tobytes()
    d["amp_sha256"] = str(d["amp_sha256"])
    return d


def oracle_run(oracle, g, df18, **kw):
    """File semantics: the reference reads ceil(n/4)*4 samples, the ragged tail being stale
    buffer content = zeros for a file shorter than one 2 MiB read (air.c:233-239)."""
    return oracle.decode(g["samples"], accept_df18=df18, **kw)


@pytest.mark.parametrize("path", GOLDEN, ids=[p.stem for p in GOLDEN])
def test_oracle_matches_reference_golden(oracle, path):
    g = load(path)
    r = oracle_run(oracle, g, False)
    assert r.text(0) == g["avr"]
    assert r.text(1) == g["avrmlat"]
    assert [list(r.tried), list(r.ok)] == g["stats"].tolist()
    ra = oracle_run(oracle, g, True)
    assert ra.text(1) == g["avrmlat_a"]
    assert ra.text(2) == g["beast_a"]
    assert [list(ra.tried), list(ra.ok)] == g["stats_a"].tolist()


@pytest.mark.parametrize("path", GOLDEN, ids=[p.stem for p in GOLDEN])
def test_oracle_envelope_matches_reference_golden(oracle, path):
    g = load(path)
    r = oracle_run(oracle, g, False, want_amp=True)
    n = int(g["amp_len"])
    assert r.amp.size >= n
    assert np.array_equal(r.amp[:4096].view(np.uint32), g["amp_head"].view(np.uint32))
    assert hashlib.sha256(r.amp[:n].tobytes()).hexdigest() == g["amp_sha256"]


def test_golden_set_is_present():
    assert {p.stem for p in GOLDEN} >= {"df17_20db", "mix_errors", "dense", "snr8", "ragged_tail"}


def test_crc_table_and_known_answer(oracle):
    """crc.h:1-34 is the byte table of generator 0xFFF409; the published DF17 frame has syndrome 0."""
    tab = oracle.crc_table()
    for b in (0, 1, 2, 0x80, 0xFF):
        r = b << 16
        for _ in range(8):
            r = ((r << 1) ^ 0xFFF409) if r & 0x800000 else (r << 1)
        assert tab[b] == r & 0xFFFFFF
    assert tab[1] == 0xFFF409 and tab[255] == 0xFA0480      # crc.h first / last non-trivial entries
    msg = bytes.fromhex("8D4840D6202CC371C32CE0576098")
    crc = 0
    for v in msg[:11]:
        crc = ((crc << 8) ^ int(tab[v ^ ((crc >> 16) & 0xFF)])) & 0xFFFFFF
    assert crc ^ int.from_bytes(msg[11:], "big") == 0


def _ref_or_skip(oracle):
    b = oracle.ref_binary()
    if b is None:
        pytest.skip("oracle/_ref not built or not runnable on this host")
    return b


def _run_ref_complete(oracle, path, flags, tcp=False):
    for _ in range(50):         # EOF race of the reference, see make_golden.py
        r = (oracle.run_ref_tcp if tcp else oracle.run_ref)(path, flags=flags)
        t, k = r.stats()
        if tcp or r.stdout.count(b"\n") == sum(k):
            return r
    raise AssertionError("reference kept dropping frames")


def test_oracle_vs_live_reference(oracle, tmp_path):
    _ref_or_skip(oracle)
    cap = capture_basic(seed=5, seconds=0.3, n_frames=120, pad=1 << 21)
    f = tmp_path / "c.bin"
    cap.tofile(f)
    for flags, df18, fmt in ((("-m",), False, 1), (("-a", "-m"), True, 1), ((), False, 0)):
        ref = _run_ref_complete(oracle, f, flags)
        r = oracle.decode(cap.samples, accept_df18=df18)
        assert r.text(fmt) == ref.stdout
        assert (r.tried, r.ok) == ref.stats()


def test_oracle_vs_live_reference_no_padding_tail(oracle, tmp_path):
    """air.c:94-99: whatever is left in ampbuff at EOF is never scanned."""
    _ref_or_skip(oracle)
    cap = capture_basic(seed=6, seconds=0.05, n_frames=40, pad=0)
    for cut in (0, 4, 40000, 123456):
        s = cap.samples[:cap.samples.size - cut]
        f = tmp_path / f"t{cut}.bin"
        s.astype("<u2").tofile(f)
        ref = _run_ref_complete(oracle, f, ("-a", "-m"))
        r = oracle.decode(s, accept_df18=True)
        assert (r.tried, r.ok) == ref.stats()
        assert r.text(1) == ref.stdout


def test_oracle_envelope_vs_probe_across_fidx_wrap(oracle, tmp_path):
    """air.c:34,65: uint32 fidx; white-box probe presets it (oracle/probe/amp_probe.c)."""
    _ref_or_skip(oracle)
    cap = capture_basic(seed=7, seconds=0.01, n_frames=4, pad=0)
    f = tmp_path / "w.bin"
    cap.tofile(f)
    for origin in (0, 2**32 - 50000, 7 * 2**32 - 10):
        amp = oracle.run_amp_probe(f, tmp_path / "a.f32", origin)
        r = oracle.decode(cap.samples, want_amp=True, origin=origin)
        assert np.array_equal(amp.view(np.uint32), r.amp[:amp.size].view(np.uint32))


def test_fma_flavour_differs(oracle):
    """FMA-contracted and uncontracted builds differ bitwise; parity follows the
    reference's own flags (-O3 -march=native => contracted)."""
    cap = capture_basic(seed=8, seconds=0.01, n_frames=4, pad=0)
    a = oracle.decode(cap.samples, want_amp=True, fma=True).amp
    b = oracle.decode(cap.samples, want_amp=True, fma=False).amp
    assert 0.2 < np.mean(a.view(np.uint32) != b.view(np.uint32)) < 0.8


def test_oracle_file_ragged_odd_tail_vs_live_reference(oracle, tmp_path):
    """A file of two 2 MiB reads plus a ragged, odd-length tail: decodeiq reads stale entries of the
    reused buffer behind the tail (air.c:59-82, :233-239)."""
    _ref_or_skip(oracle)
    cap = capture_basic(seed=9, seconds=0.12, n_frames=60, pad=0)
    raw = cap.samples.astype("<u2").tobytes()[:2 * (2 << 20) + 2 * 100001 + 1]
    f = tmp_path / "ragged.bin"
    f.write_bytes(raw)
    ref = _run_ref_complete(oracle, f, ("-a", "-m"))
    r = oracle.decode_file(f, accept_df18=True)
    assert r.text(1) == ref.stdout and (r.tried, r.ok) == ref.stats()


def test_oracle_repair_extension(oracle):
    """EXTENSION (off by default, absent from the reference): 1-/2-bit repair by syndrome.  The oracle's
    brute-force statement must give back the transmitted frame, never touch the DF bits, and leave the
    default (reference) behaviour alone."""
    from adsbdec_b200.modulator import crc24, synth_capture
    cap = synth_capture(0.3, 120, snr_db=25, seed=77, df_mix=((17, 0.7), (11, 0.3)), bit_errors=1,
                        pad=1 << 16, poisson=False)
    r0 = oracle.decode(cap.samples)
    r1 = oracle.decode(cap.samples, repair=1)
    assert all(x == 0 for x in r0.fixed)
    assert len(r1.frames) > len(r0.frames) + 20
    # everything the reference mode decodes is still there, unchanged
    k1 = {(f.n, f.data) for f in r1.frames}
    assert all((f.n, f.data) in k1 for f in r0.frames)
    n_fixed = 0
    for f, nf in zip(r1.frames, r1.fixed):
        # every accepted frame now has syndrome 0 over its (possibly corrected) bytes
        assert crc24(f.data[:-3]) == int.from_bytes(f.data[-3:], "big")
        n_fixed += nf == 1
        assert nf in (0, 1)
    assert n_fixed > 20
    # a frame with an error in the DF field never comes back
    sent = {b.frame for b in cap.bursts}
    bad_df = [b for b in cap.bursts if crc24(b.frame[:-3]) != int.from_bytes(b.frame[-3:], "big")]
    assert bad_df and len(sent) == len(cap.bursts)
    # two flipped bits need level 2, and only 112-bit frames get it
    cap2 = synth_capture(0.3, 120, snr_db=25, seed=78, df_mix=((17, 0.7), (11, 0.3)), bit_errors=2,
                         pad=1 << 16, poisson=False)
    a = oracle.decode(cap2.samples, repair=1)
    b = oracle.decode(cap2.samples, repair=2)
    assert sum(x == 2 for x in b.fixed) > 15 and not any(x == 2 for x in a.fixed)
    assert all(len(f.data) == 14 for f, x in zip(b.frames, b.fixed) if x == 2)


def test_oracle_empty_and_tiny_inputs(oracle, tmp_path):
    """air.c:233-244: an empty file ends the read loop at once; anything shorter than one deqframe flush
    (40 980 envelope samples, air.c:94) is never scanned at all."""
    import numpy as np
    assert oracle.decode(np.zeros(0, dtype="<u2")).frames == []
    f = tmp_path / "empty.bin"
    f.write_bytes(b"")
    r = oracle.decode_file(f)
    assert r.frames == [] and r.tried == (0, 0, 0) and r.ok == (0, 0, 0)
    from adsbdec_b200.modulator import synth_capture
    cap = synth_capture(0.004, 8, snr_db=25, seed=5, pad=0, poisson=False)   # 80 000 samples = 40 000 outputs
    assert cap.samples.size // 2 < 40980
    assert oracle.decode(cap.samples).frames == []
    binary = oracle.ref_binary()
    if binary is not None:
        g = tmp_path / "tiny.bin"
        cap.tofile(g)
        assert oracle.run_ref(g, binary=binary).stdout == b""


def test_address_parity_extension_of_the_oracle(oracle):
    """EXTENSION (no reference behaviour to pin it to: the reference rejects these frames).  The oracle's own
    statement of the rule, checked for what must hold by construction: off by default, every accepted
    DF4/5/20/21 frame is one that was sent, its address was announced by an earlier DF11/17/18 frame of the
    capture, and an aircraft that never announces itself is never accepted."""
    from adsbdec_b200.modulator import crc24, synth_fleet_capture
    cap = synth_fleet_capture(0.3, 400, 15, snr_db=22, seed=70, pad=1 << 16)
    plain = oracle.decode(cap.samples, accept_df18=True)
    ref = oracle.decode(cap.samples, accept_df18=True, address_parity=True)
    assert not ({f.data[0] >> 3 for f in plain.frames} & {4, 5, 20, 21})
    sent = {bytes(b.frame) for b in cap.bursts}
    announced_at = {}
    for b in cap.bursts:
        if (b.frame[0] >> 3) in (11, 17, 18):
            announced_at.setdefault(bytes(b.frame[1:4]), b.start)
    extra = [f for f in ref.frames if (f.data[0] >> 3) in (4, 5, 20, 21)]
    assert len(extra) > 60
    for f in extra:
        assert bytes(f.data) in sent
        addr = (crc24(bytes(f.data[:-3])) ^ int.from_bytes(f.data[-3:], "big")).to_bytes(3, "big")
        assert addr in announced_at and announced_at[addr] // 2 < f.n


def test_iq_input_extension_of_the_oracle(oracle):
    """EXTENSION (the reference has no I/Q path): complex baseband pairs, envelope = I*I + Q*Q.  What must hold by
    construction: read as real samples the capture decodes to nothing, read as I/Q it returns frames that were
    sent, most of them, with timestamps in order."""
    from adsbdec_b200.modulator import synth_iq_capture
    cap = synth_iq_capture(0.3, 300, snr_db=20, seed=80, df_mix=((17, 0.6), (11, 0.3), (18, 0.1)), pad=1 << 16)
    assert len(oracle.decode(cap.samples, accept_df18=True).frames) == 0
    ref = oracle.decode(cap.samples, accept_df18=True, iq=True)
    sent = {bytes(b.frame) for b in cap.bursts}
    assert len(ref.frames) > 200 and all(bytes(f.data) in sent for f in ref.frames)
    ts = [f.ts for f in ref.frames]
    assert ts == sorted(ts)
    starts = {b.start // 2 for b in cap.bursts}
    assert sum(f.n in starts or f.n + 1 in starts or f.n - 1 in starts for f in ref.frames) > 0.9 * len(ref.frames)
