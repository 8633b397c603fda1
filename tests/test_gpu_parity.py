"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle on the same
seeded inputs.  Bit-exact: frame bytes, offsets, timestamps, pw, Try/Ok counters, envelope bits."""
import numpy as np
import pytest

from helpers import DF_MIX_ALL, capture_basic, frames_key
from adsbdec_b200.modulator import synth_capture

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
