"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle on the same
seeded inputs.  Bit-exact: frame bytes, offsets, timestamps, pw, Try/Ok counters, envelope bits."""
import numpy as np
import pytest

from helpers import DF_MIX_ALL, capture_basic, frames_key
from adsbdec_b200.modulator import synth_capture, synth_fleet_capture

pytestmark = pytest.mark.gpu


def run_both(adsb, oracle, samples, *, accept_df18=False, kernel=0, origin=0, **kw):
    ref = oracle.decode(samples, accept_df18=accept_df18, origin=origin)
    d = adsb.decode(samples, accept_df18=accept_df18, kernel=kernel, start_sample=origin, **kw)
    got = d.frames
    st = d.stats()
    d.close()
    return ref, got, st


def assert_same(ref, got, st):
    assert frames_key(got) == frames_key(ref.frames)
    assert st == (ref.tried, ref.ok)


def test_envelope_bitexact(adsb, oracle):
    cap = capture_basic(seed=11, seconds=0.05, n_frames=20, pad=0)
    ref = oracle.decode(cap.samples, want_amp=True)
    amp = adsb.envelope(cap.samples)
    assert amp.size == ref.amp.size
    assert np.array_equal(amp.view(np.uint32), ref.amp.view(np.uint32))


@pytest.mark.parametrize("origin", [2**32 - 40000, 2**32 - 40004 + 8, 5 * 2**32 - 8192])
def test_envelope_fidx_wrap(adsb, oracle, origin):
    """air.c:34,65: fidx is uint32; the ring slot jumps at 2^32 samples."""
    cap = capture_basic(seed=12, seconds=0.01, n_frames=6, pad=0)
    ref = oracle.decode(cap.samples, want_amp=True, origin=origin)
    amp = adsb.envelope(cap.samples, start_sample=origin)
    assert np.array_equal(amp.view(np.uint32), ref.amp.view(np.uint32))


@pytest.mark.parametrize("kernel", [1, 0], ids=["generic", "fused"])
@pytest.mark.parametrize("df18", [False, True], ids=["noa", "a"])
def test_decode_mix(adsb, oracle, kernel, df18):
    """Config 3 shape: DF11/17/18/20/21 mix; only syndrome-0 DF11/17 (and DF18 with -a) come out."""
    cap = capture_basic(seed=21, seconds=0.5, n_frames=200)
    ref, got, st = run_both(adsb, oracle, cap.samples, accept_df18=df18, kernel=kernel)
    assert len(ref.frames) > 50
    assert_same(ref, got, st)


@pytest.mark.parametrize("kernel", [1, 0], ids=["generic", "fused"])
def test_decode_bit_errors(adsb, oracle, kernel):
    """1- and 2-bit errors: the reference has no repair (valid.c:28 errcorr is dead), they vanish."""
    for nerr in (1, 2):
        cap = synth_capture(0.3, 100, snr_db=22, seed=30 + nerr, df_mix=((17, 0.6), (11, 0.4)),
                            bit_errors=nerr, pad=1 << 16)
        ref, got, st = run_both(adsb, oracle, cap.samples, kernel=kernel)
        assert 20 < len(ref.frames) < 90
        assert_same(ref, got, st)


@pytest.mark.parametrize("snr", [6, 9, 12, 15, 20])
def test_decode_snr_sweep(adsb, oracle, snr):
    """Config 4 shape: decode count and content must match at every SNR."""
    cap = synth_capture(0.5, 150, snr_db=snr, seed=40 + snr, df_mix=((17, 0.7), (11, 0.3)), pad=1 << 16)
    ref, got, st = run_both(adsb, oracle, cap.samples)
    assert_same(ref, got, st)


def test_decode_dense_overlaps(adsb, oracle):
    """Config 2 shape: dense traffic with overlapping / garbled frames."""
    cap = synth_capture(0.5, 2500, snr_db=18, seed=50, df_mix=DF_MIX_ALL, pad=1 << 16)
    ref, got, st = run_both(adsb, oracle, cap.samples, accept_df18=True)
    assert len(ref.frames) > 300
    assert_same(ref, got, st)


@pytest.mark.parametrize("n_samples", [0, 4, 2396, 2400, 2404, 81960, 84360, 84364, 84368, 200000, 1 << 20,
                                       (1 << 20) + 4])
def test_tail_and_tiny_inputs(adsb, oracle, n_samples):
    """air.c:94-99: the tail of a capture is never scanned; tiny inputs give nothing."""
    cap = synth_capture(0.06, 40, snr_db=20, seed=60, pad=0, poisson=False)
    s = cap.samples[:n_samples]
    for kernel in (1, 0):
        ref, got, st = run_both(adsb, oracle, s, kernel=kernel)
        assert_same(ref, got, st)


@pytest.mark.parametrize("kernel", [1, 0], ids=["generic", "fused"])
def test_saturated_traffic_is_one_long_greedy_chain(adsb, oracle, kernel):
    """Short frames back to back (one every 713 offsets, each 640 long): no two candidates are a full frame
    (1200 offsets) apart, so the greedy rule (demod.c:128, :134) links the whole capture into ONE chain - the
    reference's own sequential dependency.  k_greedy hands such a chain from the thread that found it to its
    warp (32 candidates per load, the rule applied out of registers)."""
    cap = synth_capture(0.1, 1400, snr_db=22, seed=61, df_mix=((11, 1.0),), pad=1 << 16, poisson=False)
    ref = oracle.decode(cap.samples, accept_df18=True)
    d = adsb.decode(cap.samples, accept_df18=True, kernel=kernel)
    assert frames_key(d.frames) == frames_key(ref.frames)
    assert d.stats() == (ref.tried, ref.ok)
    n = np.array([f.n for f in ref.frames])
    assert len(n) > 1300 and np.diff(n).max() < 1200


def test_tail_not_emulated_decodes_more(adsb, oracle):
    cap = synth_capture(0.06, 40, snr_db=20, seed=61, pad=0, poisson=False)
    d0 = adsb.decode(cap.samples)
    d1 = adsb.decode(cap.samples, emulate_reference_tail=False)
    k0, k1 = frames_key(d0.frames), frames_key(d1.frames)
    assert k1[:len(k0)] == k0 and len(k1) >= len(k0)


@pytest.mark.parametrize("kernel", [1, 0], ids=["generic", "fused"])
def test_push_in_pieces(adsb, oracle, kernel):
    """decodeiq is called per 1 Mi-sample read (air.c:239); any split in multiples of 4 is the same stream."""
    cap = capture_basic(seed=70, seconds=0.4, n_frames=150)
    ref = oracle.decode(cap.samples)
    rng = np.random.default_rng(7)
    d = adsb.Decoder(kernel=kernel, batch_samples=1 << 20)
    pos = 0
    while pos < cap.samples.size:
        n = int(rng.integers(1, 200000)) * 4
        d.push(cap.samples[pos:pos + n])
        pos += n
    d.finish()
    assert frames_key(d.frames) == frames_key(ref.frames)
    assert d.stats() == (ref.tried, ref.ok)


@pytest.mark.parametrize("kernel", [1, 0], ids=["generic", "fused"])
def test_live_source_flush_and_poll(adsb, oracle, kernel):
    """A live source (rx_callback -> decodeiq, air.c:173-177) pushes small buffers and wants its frames
    soon: flush() scans what is staged, poll() hands out finished jobs.  Neither changes the output."""
    import time
    cap = capture_basic(seed=71, seconds=0.4, n_frames=150)
    ref = oracle.decode(cap.samples)
    rng = np.random.default_rng(8)
    seen = []
    d = adsb.Decoder(kernel=kernel, sink=lambda fr, ts, pw: seen.append((ts, pw, fr)))
    pos, early = 0, 0
    while pos < cap.samples.size:
        n = int(rng.integers(1, 60000)) * 4
        d.push(cap.samples[pos:pos + n])
        pos += n
        early += d.flush()
        if rng.integers(0, 3) == 0:
            time.sleep(0.002)
        early += d.poll()
    before_finish = len(seen)
    d.finish()
    assert frames_key(d.frames) == frames_key(ref.frames)
    assert seen == [(f.ts, f.pw, f.data) for f in ref.frames]
    assert d.stats() == (ref.tried, ref.ok)
    # most frames left long before the end of the stream (default batch = 32 Mi samples > this capture)
    assert early == before_finish and before_finish > 0.8 * len(ref.frames)
    assert d.poll() == 0
    with pytest.raises(adsb.AdsbError):
        d.flush()
    d.close()


@pytest.mark.parametrize("kernel", [1, 0], ids=["generic", "fused"])
def test_packed12_input_is_the_same_stream(adsb, oracle, kernel):
    """the airspy's packed 12-bit wire format (8 samples in 12 bytes), unpacked on the GPU into
    the same sample window.  Whole capture, pieces, and mixed with 16-bit pushes: always the oracle's output."""
    cap = capture_basic(seed=72, seconds=0.4, n_frames=150)
    ref = oracle.decode(cap.samples)
    n = cap.samples.size - cap.samples.size % 8
    words = adsb.pack12(cap.samples[:n])
    d = adsb.Decoder(kernel=kernel, batch_samples=1 << 20)
    d.push_packed12(words)
    d.push(cap.samples[n:])
    d.finish()
    assert frames_key(d.frames) == frames_key(ref.frames) and d.stats() == (ref.tried, ref.ok)
    rng = np.random.default_rng(9)
    d.reset()
    pos = 0
    while pos < n:
        k = min(int(rng.integers(1, 100000)) * 8, n - pos)
        if rng.integers(0, 2):
            d.push_packed12(words[pos // 8 * 3:(pos + k) // 8 * 3])
        else:
            d.push(cap.samples[pos:pos + k])
        pos += k
    d.push(cap.samples[n:])
    d.finish()
    assert frames_key(d.frames) == frames_key(ref.frames) and d.stats() == (ref.tried, ref.ok)
    with pytest.raises(adsb.AdsbError):
        d.reset()
        d.push_packed12(words[:3], 4)
    d.close()


def test_push_rejects_ragged_length(adsb):
    d = adsb.Decoder()
    with pytest.raises(adsb.AdsbError):
        d.push(np.zeros(6, dtype=np.uint16))


@pytest.mark.parametrize("origin", [2**32 - 3_000_000, 2**32 - 3_000_000 + 4, 3 * 2**32 - 1_000_000])
def test_decode_across_fidx_wrap(adsb, oracle, origin):
    """A stream that crosses the uint32 wrap of fidx: fused spans before/after, generic at the seam."""
    cap = capture_basic(seed=80, seconds=0.4, n_frames=300)
    ref, got, st = run_both(adsb, oracle, cap.samples, origin=origin)
    assert len(ref.frames) > 80
    assert_same(ref, got, st)


@pytest.mark.parametrize("origin", [4, 8, 12, 1000, 2**31 + 20])
def test_decode_phase_origins(adsb, oracle, origin):
    """FIR rotation phases other than 0 at the stream start (time-chunk decoders use these)."""
    cap = capture_basic(seed=81, seconds=0.2, n_frames=100)
    ref, got, st = run_both(adsb, oracle, cap.samples, origin=origin)
    assert_same(ref, got, st)


def test_wide_samples_fall_back(adsb, oracle):
    """Samples beyond 12 bits: the fused kernel's float tricks do not apply; it must notice and the
    exact generic kernels must take over."""
    cap = capture_basic(seed=90, seconds=0.1, n_frames=40)
    s = cap.samples.copy()
    s[50_000:50_400] = 30000
    s[700_001] = 4096
    ref, got, st = run_both(adsb, oracle, s)
    assert_same(ref, got, st)


def test_fused_vs_generic_large_noise(adsb):
    """Differential test of the two independent CUDA formulations on 64 Mi samples."""
    rng = np.random.default_rng(99)
    s = np.clip(np.rint(rng.standard_normal(1 << 26, dtype=np.float32) * 40 + 2048), 0, 4095).astype("<u2")
    cap = synth_capture(1.0, 3000, snr_db=15, seed=98, df_mix=DF_MIX_ALL, pad=0)
    s[:cap.samples.size] = cap.samples
    a = adsb.decode(s, kernel=0, accept_df18=True)
    b = adsb.decode(s, kernel=1, accept_df18=True)
    assert frames_key(a.frames) == frames_key(b.frames)
    assert a.stats() == b.stats()
    assert len(a.frames) > 500
    assert a.fused_time()[1] >= 1 and b.fused_time()[1] == 0


def test_text_formats(adsb, oracle):
    cap = capture_basic(seed=100, seconds=0.2, n_frames=80)
    ref = oracle.decode(cap.samples, accept_df18=True)
    d = adsb.decode(cap.samples, accept_df18=True)
    for fmt in (0, 1, 2):
        assert d.text(fmt) == ref.text(fmt)


def test_sink_order_and_content(adsb, oracle):
    cap = capture_basic(seed=101, seconds=0.2, n_frames=80)
    ref = oracle.decode(cap.samples)
    got = []
    d = adsb.Decoder(sink=lambda fr, ts, pw: got.append((ts, pw, fr)))
    d.push(cap.samples)
    d.finish()
    assert got == [(f.ts, f.pw, f.data) for f in ref.frames]
    assert [g[0] for g in got] == sorted(g[0] for g in got)


# ---- golden vectors of the reference binary, file path, command line -------------------------

import subprocess
from pathlib import Path

GOLDEN = sorted((Path(__file__).parent / "golden").glob("*.npz"))


def _load_golden(p):
    z = np.load(p)
    d = {k: z[k] for k in z.files}
    for k in ("avr", "avrmlat", "avrmlat_a", "beast_a"):
        d[k] = d[k].tobytes()
    return d


@pytest.mark.parametrize("path", GOLDEN, ids=[p.stem for p in GOLDEN])
@pytest.mark.parametrize("kernel", [0, 1], ids=["fused", "generic"])
def test_golden_reference_outputs_via_file(adsb, tmp_path, path, kernel):
    """decode_file() (fileInput, air.c:219-246) on the exact bytes the reference read, against
    what the reference binary printed (tests/golden/make_golden.py)."""
    g = _load_golden(path)
    f = tmp_path / "cap.bin"
    g["samples"].astype("<u2").tofile(f)
    d = adsb.Decoder(kernel=kernel)
    assert d.decode_file(f) == 0
    assert d.text(0) == g["avr"] and d.text(1) == g["avrmlat"]
    assert [list(x) for x in d.stats()] == g["stats"].tolist()
    da = adsb.Decoder(kernel=kernel, accept_df18=True)
    assert da.decode_file(f) == 0
    assert da.text(1) == g["avrmlat_a"] and da.text(2) == g["beast_a"]
    assert [list(x) for x in da.stats()] == g["stats_a"].tolist()


def _parse_stats(stderr: str):
    t = k = None
    for line in stderr.splitlines():
        if line.startswith("Try :"):
            t = [int(x) for x in line.split(":")[1].split()]
        elif line.startswith("Ok :"):
            k = [int(x) for x in line.split(":")[1].split()]
    return [t, k]


@pytest.mark.parametrize("path", GOLDEN, ids=[p.stem for p in GOLDEN])
def test_cli_matches_golden(adsb, tmp_path, path):
    """The adsbdec-compatible command line: same flags, same stdout bytes, same stats block."""
    g = _load_golden(path)
    f = tmp_path / "cap.bin"
    g["samples"].astype("<u2").tofile(f)
    cli = str(adsb.CLI_PATH)
    r = subprocess.run([cli, "-f", str(f)], capture_output=True)
    assert r.returncode == 0 and r.stdout == g["avr"]
    assert _parse_stats(r.stderr.decode()) == g["stats"].tolist()
    r = subprocess.run([cli, "-m", "-f", str(f)], capture_output=True)
    assert r.stdout == g["avrmlat"]
    r = subprocess.run([cli, "-b", "-a", "-m", "-f", str(f)], capture_output=True)   # last of -m/-b wins (main.c:79-84)
    assert r.stdout == g["avrmlat_a"]
    assert _parse_stats(r.stderr.decode()) == g["stats_a"].tolist()
    r = subprocess.run([cli, "-a", "-b", "-f", str(f)], capture_output=True)         # full beast packets on stdout
    assert r.stdout == g["beast_a"]


@pytest.mark.parametrize("path", GOLDEN, ids=[p.stem for p in GOLDEN])
def test_patched_reference_binary_matches_golden(adsb, oracle, tmp_path, path):
    """INTEGRATION.md section 3, compiled: the reference's OWN main.c / output.c / demod.c with
    integration/apply_patch.py applied to air.c (fileInput -> libadsbdec_b200) and valid.c (stats setter),
    built by `make -C oracle patched`.  Same stdout, same beast bytes over TCP, same stats block as the
    unmodified reference printed for the golden captures: the C ABI is a drop-in for air.c:219-246 and
    feeds output.c:159 (netout) like validShort/validLong did."""
    patched = oracle.REF_DIR / "adsbdec_patched"
    if not patched.exists():
        if not any(oracle.REF_DIR.glob("adsbdec_*")):
            pytest.skip("oracle/_ref was never built in this tree")
        pytest.fail("oracle/_ref/adsbdec_patched is missing: `make -C oracle patched` (needs the built library)")
    g = _load_golden(path)
    f = tmp_path / "cap.bin"
    g["samples"].astype("<u2").tofile(f)
    r = subprocess.run([str(patched), "-f", str(f)], capture_output=True, timeout=120)
    assert r.returncode == 0 and r.stdout == g["avr"], r.stderr[-300:]
    assert _parse_stats(r.stderr.decode()) == g["stats"].tolist()
    r = subprocess.run([str(patched), "-a", "-m", "-f", str(f)], capture_output=True, timeout=120)
    assert r.stdout == g["avrmlat_a"]
    assert _parse_stats(r.stderr.decode()) == g["stats_a"].tolist()
    rb = oracle.run_ref_tcp(f, flags=("-a", "-b"), binary=patched, timeout=120)
    assert rb.stdout == g["beast_a"]


def test_cli_beast_over_tcp_and_missing_file(adsb, tmp_path):
    import socket
    import threading
    g = _load_golden(GOLDEN[0])
    f = tmp_path / "cap.bin"
    g["samples"].astype("<u2").tofile(f)
    srv = socket.socket()
    srv.bind(("127.0.0.1", 0))
    srv.listen(1)
    port = srv.getsockname()[1]
    got = []

    def sink():
        c, _ = srv.accept()
        with c:
            while True:
                b = c.recv(65536)
                if not b:
                    break
                got.append(b)
    th = threading.Thread(target=sink, daemon=True)
    th.start()
    r = subprocess.run([str(adsb.CLI_PATH), "-a", "-b", "-s", f"127.0.0.1:{port}", "-f", str(f)], capture_output=True)
    th.join(10)
    assert r.returncode == 0 and b"".join(got) == g["beast_a"]
    # air.c:225-228: a file that cannot be opened -> no frames, stats, exit 0
    r = subprocess.run([str(adsb.CLI_PATH), "-f", str(tmp_path / "nope.bin")], capture_output=True)
    assert r.returncode == 0 and r.stdout == b"" and b"Try :" in r.stderr


def test_cli_listener_serves_the_first_client(adsb, tmp_path):
    """-l addr:port (main.c:69-72, output.c:130-146): bind, print "listening", accept one client, then
    write to it like -s.  Same bytes as the reference's beast output over TCP."""
    import socket
    import time
    g = _load_golden(GOLDEN[0])
    f = tmp_path / "cap.bin"
    g["samples"].astype("<u2").tofile(f)
    probe = socket.socket()
    probe.bind(("127.0.0.1", 0))
    port = probe.getsockname()[1]
    probe.close()
    p = subprocess.Popen([str(adsb.CLI_PATH), "-a", "-b", "-l", f"127.0.0.1:{port}", "-f", str(f)],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    try:
        assert p.stderr.readline().strip() == b"listening"
        got = b""
        for _ in range(50):
            try:
                c = socket.create_connection(("127.0.0.1", port), timeout=5)
                break
            except OSError:
                time.sleep(0.1)
        with c:
            while True:
                b = c.recv(65536)
                if not b:
                    break
                got += b
        out, err = p.communicate(timeout=60)
    finally:
        if p.poll() is None:
            p.kill()
    assert p.returncode == 0 and out == b"" and got == g["beast_a"]
    assert err.startswith(b"connected") and b"Try :" in err


def test_cli_vs_live_reference_binary(adsb, oracle, tmp_path):
    """When oracle/_ref travelled to this box and runs here: byte-identical stdout on a 1 s capture."""
    binary = oracle.ref_binary()
    if binary is None:
        pytest.skip("oracle/_ref not present / not runnable on this host")
    cap = synth_capture(1.0, 400, snr_db=18, seed=123, df_mix=DF_MIX_ALL, pad=1 << 21)
    f = tmp_path / "c1.bin"
    cap.tofile(f)
    for flags in (["-m"], ["-a", "-m"], []):
        for _ in range(30):     # reference EOF race: retry until nothing was dropped
            ref = oracle.run_ref(f, flags=tuple(flags), binary=binary)
            t, k = ref.stats()
            if ref.stdout.count(b"\n") == sum(k):
                break
        r = subprocess.run([str(adsb.CLI_PATH), *flags, "-f", str(f)], capture_output=True)
        assert r.stdout == ref.stdout
        assert _parse_stats(r.stderr.decode()) == [list(t), list(k)]


# ---- device-resident input, time-chunking, full-size properties ------------------------------

def test_push_device_matches_host_push(adsb, oracle):
    import torch
    cap = capture_basic(seed=140, seconds=0.6, n_frames=300)
    ref = oracle.decode(cap.samples, accept_df18=True)
    t = torch.from_numpy(cap.samples.astype(np.int16)).cuda()
    d = adsb.Decoder(accept_df18=True)
    d.push_device(t.data_ptr(), t.numel())
    d.finish()
    assert frames_key(d.frames) == frames_key(ref.frames) and d.stats() == (ref.tried, ref.ok)
    # in pieces: aligned big ones take the direct path, small / ragged ones the staging window
    d2 = adsb.Decoder(accept_df18=True, batch_samples=1 << 20)
    pos = 0
    for n in (400_000, 8, 2_000_000, 4, 1_048_576, 12):
        d2.push_device(t.data_ptr() + 2 * pos, n)
        pos += n
    d2.push(cap.samples[pos:pos + 1000])
    pos += 1000
    d2.push_device(t.data_ptr() + 2 * pos, t.numel() - pos)
    d2.finish()
    assert frames_key(d2.frames) == frames_key(ref.frames) and d2.stats() == (ref.tried, ref.ok)


@pytest.mark.parametrize("n_chunks", [2, 5])
def test_chunk_candidates_merge_equals_whole(adsb, oracle, n_chunks):
    """time chunks with halo scanned independently (any GPU), merged by offset, resolved once."""
    import torch
    from adsbdec_b200.chunking import merge_candidates, plan_chunks, resolve_merged
    cap = capture_basic(seed=141, seconds=0.5, n_frames=400)
    ref = oracle.decode(cap.samples, accept_df18=True)
    t = torch.from_numpy(cap.samples.astype(np.int16)).cuda()
    d = adsb.Decoder(accept_df18=True)
    parts = []
    for ch in plan_chunks(cap.samples.size, n_chunks):
        seg = t[ch.sample_begin:ch.sample_end].clone()       # a chunk as it would sit on another GPU
        lo, hi = ch.local_own
        c, _tries = d.chunk_candidates(seg.data_ptr(), seg.numel(), ch.sample_begin, lo, hi)
        parts.append((ch, c))
    out = resolve_merged(merge_candidates(parts), cap.samples.size)
    assert [(f.ts, f.pw, f.data) for f in out] == [(f.ts, f.pw, f.data) for f in ref.frames]


def test_full_size_60s_tiling_property(adsb, oracle):
    """BASELINE configs[1] size (1.2 G samples, 2.4 GB in HBM).  The capture is one 2 s dense tile
    repeated 30 times; the tile length is a multiple of 28 samples (FIR rotation period 14, mixer
    period 4) and ends in silence, so every tile must decode to the oracle's frames of ONE tile,
    shifted: same bytes and pw, offsets + k*tile, timestamps following the drift rule."""
    import torch
    tile_cap = synth_capture(2.0, 6000, snr_db=16, seed=150, df_mix=DF_MIX_ALL, pad=0)
    s = tile_cap.samples
    s = s[:s.size - s.size % 28].copy()
    s[-6000:] = 2048
    ref = oracle.decode(np.concatenate([s, np.full(1 << 17, 2048, dtype="<u2")]), accept_df18=True)
    assert len(ref.frames) > 1500
    reps = 30
    tile = torch.from_numpy(s.astype(np.int16)).cuda()
    big = torch.cat([tile.repeat(reps), torch.full((1 << 17,), 2048, dtype=torch.int16, device="cuda")])
    assert big.numel() >= 1_199_000_000
    d = adsb.Decoder(accept_df18=True)
    d.push_device(big.data_ptr(), big.numel())
    d.finish()
    fa = d.frames_array
    nt = len(ref.frames)
    assert len(fa) == reps * nt
    tile_out = s.size // 2
    ref_n = np.array([f.n for f in ref.frames], dtype=np.uint64)
    ref_b = np.array([np.frombuffer(f.data.ljust(14, b"\0"), dtype=np.uint8) for f in ref.frames])
    ref_pw = np.array([f.pw for f in ref.frames], dtype=np.uint32)
    skip = np.array([640 if len(f.data) == 7 else 1200 for f in ref.frames], dtype=np.uint64)
    drift_tile = int((skip - 1).sum())
    drift_before = np.concatenate([[0], np.cumsum(skip - 1)[:-1]]).astype(np.uint64)
    for k in (0, 1, 7, reps - 1):
        blk = fa[k * nt:(k + 1) * nt]
        assert np.array_equal(blk["n"], ref_n + np.uint64(k * tile_out))
        assert np.array_equal(blk["bytes"], ref_b) and np.array_equal(blk["pw"], ref_pw)
        ts = ref_n + np.uint64(k * tile_out + 1) - drift_before - np.uint64(k * drift_tile)
        assert np.array_equal(blk["ts"], ts)
    # and the whole list: periodic bytes, strictly increasing timestamps
    assert np.array_equal(fa["bytes"].reshape(reps, nt, 14), np.broadcast_to(ref_b, (reps, nt, 14)))
    assert np.all(np.diff(fa["ts"].astype(np.int64)) > 0)
    tr, ok = d.stats()
    assert ok == tuple(reps * x for x in ref.ok) and tr == tuple(reps * x for x in ref.tried)


@pytest.mark.parametrize("kernel", [0, 1], ids=["fused", "generic"])
def test_list_overflow_regrows_and_redoes_in_flight_jobs(adsb, oracle, kernel):
    """Tiny initial candidate / try lists on a multi-batch stream: every job overflows, the decoder
    must rewind the device-side scan position, regrow and re-issue the jobs in flight - and still
    produce exactly the oracle's output."""
    cap = synth_capture(0.8, 900, snr_db=20, seed=160, df_mix=DF_MIX_ALL, pad=1 << 16)
    ref = oracle.decode(cap.samples, accept_df18=True)
    d = adsb.Decoder(accept_df18=True, kernel=kernel, batch_samples=1 << 21, initial_caps=(16, 64))
    d.push(cap.samples)
    d.finish()
    assert frames_key(d.frames) == frames_key(ref.frames)
    assert d.stats() == (ref.tried, ref.ok)
    assert len(ref.frames) > 300


@pytest.mark.parametrize("extra_bytes", [0, 2 * 100001 + 1, 2 * 6, 2 * 4 + 1])
def test_decode_file_multi_read_ragged_tail(adsb, oracle, tmp_path, extra_bytes):
    """decode_file reads straight into page-locked staging; the stale entries decodeiq would see
    behind a ragged / odd tail (air.c:59-82) are fetched from the previous read's file positions."""
    cap = capture_basic(seed=170, seconds=0.12, n_frames=60, pad=0)
    raw = cap.samples.astype("<u2").tobytes()[:2 * (2 << 20) + extra_bytes]
    f = tmp_path / "ragged.bin"
    f.write_bytes(raw)
    ref = oracle.decode_file(f, accept_df18=True)
    d = adsb.Decoder(accept_df18=True)
    assert d.decode_file(f) == 0
    assert frames_key(d.frames) == frames_key(ref.frames) and d.stats() == (ref.tried, ref.ok)
    r = subprocess.run([str(adsb.CLI_PATH), "-a", "-m", "-f", str(f)], capture_output=True)
    assert r.stdout == ref.text(1)


# ---- handle life cycle and error behaviour ---------------------------------------------------

def test_reset_reuses_the_handle(adsb, oracle):
    a = capture_basic(seed=180, seconds=0.2, n_frames=80)
    b = capture_basic(seed=181, seconds=0.15, n_frames=60)
    d = adsb.Decoder(accept_df18=True)
    for cap, origin in ((a, 0), (b, 0), (a, 2**32 - 1_000_000), (b, 8)):
        d.reset(origin)
        d.push(cap.samples)
        d.finish()
        ref = oracle.decode(cap.samples, accept_df18=True, origin=origin)
        assert frames_key(d.frames) == frames_key(ref.frames) and d.stats() == (ref.tried, ref.ok)


def test_state_errors(adsb):
    d = adsb.Decoder()
    d.push(np.full(4000, 2048, dtype="<u2"))
    d.finish()
    d.finish()                                   # idempotent
    with pytest.raises(adsb.AdsbError) as e:
        d.push(np.full(8, 2048, dtype="<u2"))    # push after finish
    assert e.value.code == -5
    with pytest.raises(adsb.AdsbError) as e:
        adsb.Decoder(start_sample=2)             # must be a multiple of 4
    assert e.value.code == -1
    with pytest.raises(adsb.AdsbError) as e:
        adsb.Decoder(device=99)
    assert e.value.code == -2
    with pytest.raises(adsb.AdsbError) as e:
        d.decode_file("/nonexistent/file.bin")   # finished handle
    assert e.value.code == -5
    assert adsb.Decoder().decode_file("/nonexistent/file.bin") == -6   # air.c:225-228: no frames, no error exit


def test_stats_can_be_switched_off(adsb, oracle):
    cap = capture_basic(seed=182, seconds=0.2, n_frames=80)
    ref = oracle.decode(cap.samples)
    d = adsb.decode(cap.samples, collect_stats=False)
    assert frames_key(d.frames) == frames_key(ref.frames)
    tr, ok = d.stats()
    assert tr == (0, 0, 0) and ok == ref.ok


def test_clear_frames_on_a_long_running_stream(adsb, oracle):
    cap = capture_basic(seed=183, seconds=0.5, n_frames=250)
    ref = oracle.decode(cap.samples)
    got = []
    d = adsb.Decoder(batch_samples=1 << 20, sink=lambda fr, ts, pw: got.append((ts, pw, fr)))
    for i in range(0, cap.samples.size, 1 << 20):
        d.push(cap.samples[i:i + (1 << 20)])
        adsb.lib().adsbdec_b200_clear_frames(d._h)
    d.finish()
    assert got == [(f.ts, f.pw, f.data) for f in ref.frames]
    assert d.stats() == (ref.tried, ref.ok)
    assert d.num_frames <= len(ref.frames)


def test_many_concurrent_receiver_streams(adsb, oracle):
    """Config 5 shape: many independent receiver streams decoded side by side on one GPU, their
    pushes interleaved; every stream must equal its own oracle run."""
    import torch
    n_streams = 24
    caps = [synth_capture(0.15, 50 + 3 * i, snr_db=14 + (i % 8), seed=300 + i, df_mix=DF_MIX_ALL, pad=1 << 15)
            for i in range(n_streams)]
    refs = [oracle.decode(c.samples, accept_df18=bool(i & 1)) for i, c in enumerate(caps)]
    decs = [adsb.Decoder(accept_df18=bool(i & 1), batch_samples=1 << 19) for i in range(n_streams)]
    free0 = torch.cuda.mem_get_info()[0]
    piece = 262144
    for pos in range(0, max(c.samples.size for c in caps), piece):
        for d, c in zip(decs, caps):
            if pos < c.samples.size:
                d.push(c.samples[pos:pos + piece])
    for d in decs:
        d.finish()
    used = free0 - torch.cuda.mem_get_info()[0]
    for d, r in zip(decs, refs):
        assert frames_key(d.frames) == frames_key(r.frames) and d.stats() == (r.tried, r.ok)
    assert used / n_streams < 64 << 20, f"{used / n_streams / 2**20:.0f} MiB of HBM per stream"


# (the stream that crosses a REAL 2^32 fidx wrap is decoded against the live reference binary in
# tests/test_gpu_at_size.py::test_config5_stream_across_a_real_fidx_wrap)


@pytest.mark.parametrize("kernel", [1, 0], ids=["generic", "fused"])
@pytest.mark.parametrize("level", [1, 2])
def test_repair_extension_matches_oracle(adsb, oracle, kernel, level):
    """EXTENSION, off by default (the reference has no repair): syndrome-lookup repair of one / two flipped
    bits on the GPU against the oracle's brute-force statement of the same rule - frames, corrected bytes,
    number of bits corrected, timestamps (a repaired frame takes part in the greedy skip) and Try/Ok."""
    for nerr in (1, 2):
        cap = synth_capture(0.3, 150, snr_db=22, seed=40 + nerr, df_mix=((17, 0.5), (18, 0.2), (11, 0.3)),
                            bit_errors=nerr, pad=1 << 16)
        ref = oracle.decode(cap.samples, accept_df18=True, repair=level)
        d = adsb.decode(cap.samples, accept_df18=True, kernel=kernel, repair=level)
        got, st, fixed = d.frames, d.stats(), [int(x) for x in d.frames_array["pad"][:, 0]]
        d.close()
        assert_same(ref, got, st)
        assert fixed == ref.fixed
        if level >= nerr:   # the injected errors are within reach of this level
            plain = oracle.decode(cap.samples, accept_df18=True)
            assert len(ref.frames) > len(plain.frames) + 10
            assert nerr in fixed


def test_repair_is_off_by_default(adsb, oracle):
    cap = synth_capture(0.2, 80, snr_db=22, seed=45, bit_errors=1, pad=1 << 16)
    d = adsb.decode(cap.samples)
    assert not d.frames_array["pad"].any()
    assert frames_key(d.frames) == frames_key(oracle.decode(cap.samples).frames)
    d.close()
    with pytest.raises(adsb.AdsbError):
        adsb.Decoder(repair=3)


@pytest.mark.parametrize("kernel", [1, 0], ids=["generic", "fused"])
@pytest.mark.parametrize("batch", [0, 1 << 20], ids=["one_job", "many_jobs"])
def test_address_parity_extension_matches_oracle(adsb, oracle, kernel, batch):
    """EXTENSION, off by default (the reference rejects every frame with a non-zero syndrome): DF4 / DF5 / DF20 /
    DF21 carry CRC ^ ICAO address in their parity field and are accepted when a clean DF11 / DF17 / DF18 frame has
    announced that address at an earlier offset.  The GPU (address table updated job by job, a gate and a slicer
    pass of their own) against the oracle's sequential statement of the same rule: frames, timestamps (an accepted
    frame takes part in the greedy skip), Try/Ok (which do not count the new classes)."""
    cap = synth_fleet_capture(0.3, 400, 15, snr_db=22, seed=70, pad=1 << 16)
    ref = oracle.decode(cap.samples, accept_df18=True, address_parity=True)
    d = adsb.decode(cap.samples, accept_df18=True, kernel=kernel, address_parity=True, batch_samples=batch)
    got, st = d.frames, d.stats()
    d.close()
    assert_same(ref, got, st)
    plain = oracle.decode(cap.samples, accept_df18=True)
    dfs = [f.data[0] >> 3 for f in ref.frames]
    assert len(ref.frames) > len(plain.frames) + 60 and {4, 5, 20, 21} <= set(dfs)
    # nobody is accepted before being announced, strangers never
    first = {}
    for f in ref.frames:
        df = f.data[0] >> 3
        if df in (11, 17, 18):
            first.setdefault(bytes(f.data[1:4]), f.n)
    sent = {bytes(b.frame): b for b in cap.bursts}
    for f in ref.frames:
        if (f.data[0] >> 3) in (4, 5, 20, 21):
            assert bytes(f.data) in sent


def test_cli_address_parity_flag(adsb, oracle, tmp_path):
    """-P (extension; the reference has no such flag and rejects these frames): same bytes as the oracle's text."""
    cap = synth_fleet_capture(0.2, 250, 10, snr_db=22, seed=72, pad=1 << 21)
    f = tmp_path / "fleet.bin"
    cap.samples.astype("<u2").tofile(f)
    r = subprocess.run([str(adsb.CLI_PATH), "-a", "-P", "-m", "-f", str(f)], capture_output=True)
    assert r.returncode == 0
    assert r.stdout == oracle.decode(cap.samples, accept_df18=True, address_parity=True).text(1)
    r = subprocess.run([str(adsb.CLI_PATH), "-a", "-m", "-f", str(f)], capture_output=True)
    assert r.stdout == oracle.decode(cap.samples, accept_df18=True).text(1)


def test_address_parity_is_off_by_default(adsb, oracle):
    cap = synth_fleet_capture(0.2, 200, 10, snr_db=22, seed=71, pad=1 << 16)
    d = adsb.decode(cap.samples, accept_df18=True)
    ref = oracle.decode(cap.samples, accept_df18=True)
    assert_same(ref, d.frames, d.stats())
    assert not ({f.data[0] >> 3 for f in d.frames} & {4, 5, 20, 21})
    d.close()
    with pytest.raises(adsb.AdsbError):
        adsb.Decoder(address_parity=True).chunk_candidates(16, 1 << 16, 0, 0, 100)


def test_frames_view_is_the_store_itself(adsb, oracle):
    """Decoder.frames_view hands out adsbdec_b200_frames() without a copy (what bench.py's end-to-end leg reads)."""
    cap = capture_basic(seed=73)
    d = adsb.decode(cap.samples, accept_df18=True)
    v, a = d.frames_view, d.frames_array
    assert len(v) == d.num_frames == len(oracle.decode(cap.samples, accept_df18=True).frames)
    assert v.tobytes() == a.tobytes() and not v.flags.owndata
    d.close()


@pytest.mark.parametrize("batch", [0, 1 << 20], ids=["one_job", "many_jobs"])
def test_iq_input_extension_matches_oracle(adsb, oracle, batch, tmp_path):
    """EXTENSION, off by default (the reference only ever asks the airspy for real samples, air.c:128-152): the
    stream is complex baseband, signed 16-bit I/Q pairs at 10 MS/s, one pair per envelope sample, envelope =
    I*I + Q*Q; demod.c / valid.c behind it as always.  Against the oracle's statement of the same: frames,
    timestamps, pw, Try/Ok - alone, with the address/parity extension, and through the command line (-I)."""
    from adsbdec_b200.modulator import synth_iq_capture
    cap = synth_iq_capture(0.3, 300, snr_db=20, seed=80, df_mix=((17, 0.6), (11, 0.3), (18, 0.1)), pad=1 << 16)
    ref = oracle.decode(cap.samples, accept_df18=True, iq=True)
    d = adsb.decode(cap.samples, accept_df18=True, iq=True, batch_samples=batch)
    assert_same(ref, d.frames, d.stats())
    d.close()
    sent = {bytes(b.frame) for b in cap.bursts}
    assert len(ref.frames) > 200 and all(bytes(f.data) in sent for f in ref.frames)
    fleet = synth_iq_capture(0.3, 400, snr_db=20, seed=81, fleet=12, pad=1 << 16)
    ref = oracle.decode(fleet.samples, accept_df18=True, iq=True, address_parity=True)
    d = adsb.decode(fleet.samples, accept_df18=True, iq=True, address_parity=True, batch_samples=batch)
    assert_same(ref, d.frames, d.stats())
    d.close()
    assert {4, 5, 20, 21} <= {f.data[0] >> 3 for f in ref.frames}
    with pytest.raises(adsb.AdsbError):      # packed 12-bit samples are real samples
        adsb.Decoder(iq=True).push_packed12(np.zeros(12, np.uint32), 32)
    if batch == 0:
        f = tmp_path / "iq.bin"
        np.concatenate([cap.samples, np.zeros(1 << 21, np.uint16)]).astype("<u2").tofile(f)
        r = subprocess.run([str(adsb.CLI_PATH), "-a", "-I", "-m", "-f", str(f)], capture_output=True)
        want = oracle.decode(np.concatenate([cap.samples, np.zeros(1 << 21, np.uint16)]), accept_df18=True, iq=True)
        assert r.returncode == 0 and r.stdout == want.text(1)
