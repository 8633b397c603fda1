"""CPU: the host side of the product - C-ABI exports, formatpkt restatement, the sequential
resolver, chunk planning/merging (incl. a world_size-2 gloo run), loud failure without a GPU."""
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from helpers import all_offset_candidates, capture_basic, frames_key

ROOT = Path(__file__).resolve().parent.parent


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_abi_exports_every_declared_symbol(adsb):
    hdr = (ROOT / "include" / "adsbdec_b200.h").read_text()
    declared = set(re.findall(r"\b(adsbdec_b200_[a-z0-9_]+)\s*\(", hdr)) - {"adsbdec_b200_sink_fn"}
    assert declared == set(adsb.EXPORTS), declared ^ set(adsb.EXPORTS)
    lib = adsb.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", str(adsb.LIB_PATH)], capture_output=True, text=True).stdout
    for name in declared:
        assert f" T {name}" in out, name


def test_header_cites_reference_lines():
    hdr = (ROOT / "include" / "adsbdec_b200.h").read_text()
    for cite in ("air.c:54", "output.c:159", "valid.c:30-31", "output.c:204-262", "air.c:219-246", "demod.c:26"):
        assert cite in hdr


def test_product_never_touches_the_oracle():
    for p in list((ROOT / "adsbdec_b200").rglob("*.py")) + list((ROOT / "adsbdec_b200" / "csrc").glob("*")):
        if p.is_file() and p.suffix in (".py", ".cu", ".cpp", ".h", ""):
            txt = p.read_text(errors="ignore")
            assert "pyoracle" not in txt and "liboracle" not in txt and "oracle/" not in txt, p


@pytest.mark.skipif(have_gpu(), reason="box has a GPU")
def test_no_gpu_fails_loudly(adsb):
    with pytest.raises(adsb.AdsbError) as e:
        adsb.Decoder()
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(adsb.AdsbError):
        adsb.envelope(np.zeros(64, dtype=np.uint16))


def test_cli_usage_and_flags(adsb):
    cli = str(adsb.CLI_PATH)
    for bad in (["-e"], ["-c"], ["-x"]):
        r = subprocess.run([cli, *bad], capture_output=True, text=True)
        assert r.returncode == 1 and "usage" in r.stdout       # main.c:85-87
    r = subprocess.run([cli], capture_output=True, text=True)
    assert r.returncode == 1
    if not have_gpu():
        r = subprocess.run([cli, "-f", "/nonexistent"], capture_output=True, text=True)
        assert r.returncode == 2 and "no CPU fallback" in r.stderr


def test_format_packet_matches_oracle(adsb, oracle):
    rng = np.random.default_rng(3)
    for _ in range(300):
        ln = 7 if rng.random() < 0.4 else 14
        fr = bytes(rng.integers(0, 256, ln, dtype=np.uint8))
        if rng.random() < 0.3:
            fr = fr[:3] + b"\x1a" + fr[4:]
        ts = int(rng.integers(0, 1 << 47))
        if rng.random() < 0.2:
            ts = (0x1A << int(rng.integers(0, 40))) * 10 // 12
        pw = int(rng.integers(0, 1 << 22))
        f = oracle.Frame(0, ts, pw, fr)
        for fmt in (0, 1, 2):
            assert adsb.format_packet(fmt, fr, ts, pw) == oracle.format_frame(f, fmt)


def test_format_packet_golden():
    import adsbdec_b200 as A
    fr = bytes.fromhex("8D4840D6202CC371C32CE0576098")
    assert A.format_packet(0, fr, 1, 0) == b"*8D4840D6202CC371C32CE0576098;\n"
    assert A.format_packet(1, fr, 10, 0) == b"@00000000000C8D4840D6202CC371C32CE0576098;\n"   # 10*12/10
    b = A.format_packet(2, fr, 0x1A * 10 // 12 + 1, 40000)          # ts12M low byte 0x1a -> doubled
    assert b[:2] == b"\x1a3" and b.count(b"\x1a") >= 1 and b.endswith(fr)
    assert A.format_packet(2, fr[:7], 0, 1600 * 1600)[8] == 1600 // 8 // 5     # output.c:215-217


@pytest.mark.parametrize("df18", [False, True])
@pytest.mark.parametrize("cut", [0, 3 * 4, 50_000, 81_000])
def test_resolver_reproduces_sequential_scan(adsb, oracle, df18, cut):
    """All-offsets candidates (numpy test double on the oracle's envelope) + the C resolver ==
    the oracle's sequential scan: same frames, offsets, timestamps; incl. the unscanned tail."""
    cap = capture_basic(seed=31, seconds=0.05, n_frames=70, pad=0)
    s = cap.samples[:cap.samples.size - cut]
    ref = oracle.decode(s, accept_df18=df18, want_amp=True)
    cands, tries = all_offset_candidates(ref.amp, df18)
    total_out = ref.amp.size
    # one shot
    r = adsb.Resolver(True)
    r.feed(cands, max(0, total_out - 1200), total_out)
    r.finish()
    assert [(f.ts, f.pw, f.data) for f in r.out] == [(f.ts, f.pw, f.data) for f in ref.frames]
    assert len(ref.frames) > 15
    # fed in pieces, as batches arrive
    r2 = adsb.Resolver(True)
    step = 30_000
    for upto in range(step, total_out + step, step):
        upto = min(upto, total_out)
        scanned = max(0, upto - 1200)
        part = cands[(cands["n"] < scanned) & (cands["n"] >= max(0, upto - step - 1200))]
        r2.feed(part, scanned, upto)
    r2.finish()
    assert [(f.ts, f.pw, f.data) for f in r2.out] == [(f.ts, f.pw, f.data) for f in ref.frames]
    # Try counters: gate-passing offsets the scan visits (valid.c:46,68)
    scan_end = r.scan_end
    spans = [(f.n, f.n + (640 if len(f.data) == 7 else 1200)) for f in ref.frames]
    vis = [0, 0, 0]
    j = 0
    for n, cls in tries:
        if n >= scan_end:
            continue
        while j < len(spans) and spans[j][1] <= n:
            j += 1
        if j < len(spans) and spans[j][0] < n < spans[j][1]:
            continue
        vis[cls] += 1
    assert tuple(vis) == ref.tried


@pytest.mark.parametrize("cut", [0, 40_000, 81_236])
def test_resolver_stamped_frames_path(adsb, oracle, cut):
    """What the stream decoder does in production: the GPU applies the greedy rule and the timestamps
    (kernels_resolve.cu; stated here in a few lines of Python over the all-offsets candidates) and the host
    resolver only runs the deqframe call schedule over the stamped frames.  Result == the oracle's sequential
    scan, tail cut-off and Ok counters included, fed whole or in pieces."""
    cap = capture_basic(seed=33, seconds=0.05, n_frames=70, pad=0)
    s = cap.samples[:cap.samples.size - cut]
    ref = oracle.decode(s, accept_df18=True, want_amp=True)
    cands, _ = all_offset_candidates(ref.amp, True)
    total_out = ref.amp.size
    stamped = np.zeros(len(cands), dtype=adsb.FRAME_DTYPE)
    k, pos, drift = 0, 0, 0
    for c in cands:                       # demod.c:99, :128, :134 over the sparse list
        if c["n"] < pos or c["n"] >= total_out - 1200:
            continue
        skip = 640 if c["len"] == 7 else 1200
        stamped[k] = (int(c["n"]) + 1 - drift, int(c["n"]), int(c["pw"]), int(c["len"]), c["bytes"][:14], (0, 0))
        k, pos, drift = k + 1, int(c["n"]) + skip, drift + skip - 1
    stamped = stamped[:k]
    want = [(f.ts, f.pw, f.data) for f in ref.frames]
    r = adsb.Resolver(True)
    ok = r.feed_frames(stamped, max(0, total_out - 1200), total_out)
    r.finish()
    assert [(f.ts, f.pw, f.data) for f in r.out] == want and len(want) > 15
    r2 = adsb.Resolver(True)
    ok2 = np.zeros(3, dtype=np.int64)
    step = 30_000
    for upto in range(step, total_out + step, step):
        upto = min(upto, total_out)
        scanned = max(0, upto - 1200)
        part = stamped[(stamped["n"] < scanned) & (stamped["n"] >= max(0, upto - step - 1200))]
        ok2 += r2.feed_frames(part, scanned, upto)
    r2.finish()
    assert [(f.ts, f.pw, f.data) for f in r2.out] == want
    assert len(stamped) >= len(want)      # the frames behind the last deqframe call were dropped
    if cut == 0:
        assert tuple(ok) == ref.ok == tuple(int(x) for x in ok2)


def test_resolver_in_place_schedule_on_keys(adsb, oracle):
    """The form the stream decoder runs in production (api.cu complete_job -> Resolver::advance_keys): all stamped
    frames stay in one array, the call schedule only sees 8-byte keys (start << 1 | long) and returns how many
    frames it has reached.  Same answer as feeding the frames themselves, job by job and at the end of the stream;
    dense traffic, so that both the short walk and the search behind it run."""
    from adsbdec_b200.modulator import synth_capture
    cap = synth_capture(0.12, 1500, snr_db=22, seed=34, df_mix=((11, 0.6), (17, 0.4)), pad=0, poisson=False)
    ref = oracle.decode(cap.samples, accept_df18=True, want_amp=True)
    cands, _ = all_offset_candidates(ref.amp, True)
    total_out = ref.amp.size
    stamped, pos, drift = [], 0, 0
    for c in cands:                       # demod.c:99, :128, :134 over the sparse list
        if c["n"] < pos or c["n"] >= total_out - 1200:
            continue
        skip = 640 if c["len"] == 7 else 1200
        stamped.append((int(c["n"]) + 1 - drift, int(c["n"]), int(c["pw"]), int(c["len"]), c["bytes"][:14], (0, 0)))
        pos, drift = int(c["n"]) + skip, drift + skip - 1
    stamped = np.array(stamped, dtype=adsb.FRAME_DTYPE)
    keys = (stamped["n"].astype(np.uint64) << np.uint64(1)) | (stamped["len"] == 14).astype(np.uint64)
    a, b = adsb.Resolver(True), adsb.Resolver(True)
    n_vis, fed, step = 0, 0, 250_000      # "jobs" of 250 k offsets
    for upto in range(step, total_out + step, step):
        upto = min(upto, total_out)
        scanned = max(0, upto - 1200)
        have = int(np.searchsorted(stamped["n"], scanned))      # frames of the jobs so far
        a.feed_frames(stamped[fed:have], scanned, upto)
        fed = have
        n_vis = b.advance_keys(keys[:have], n_vis, scanned, upto)
        assert n_vis == len(a.out) and b.scan_end == a.scan_end
    a.finish()
    n_vis = b.advance_keys(keys[:fed], n_vis, 0, 0, finish=True)
    assert n_vis == len(a.out) == len(ref.frames) > 500
    assert [(f.ts, f.data) for f in a.out] == [(f.ts, f.data) for f in ref.frames]
    assert [int(x) for x in stamped["ts"][:n_vis]] == [f.ts for f in ref.frames]


@pytest.mark.parametrize("seed", range(6))
def test_resolver_schedule_against_a_python_statement(adsb, seed):
    """air.c:94-99 as three lines of Python against Resolver::advance_keys: deqframe is called at the first even
    sample count that leaves 40 980 envelope samples queued, scans up to 1200 before the end of what it was given and
    returns where it stopped - behind a frame if one straddles that point.  Random frame lists from empty to
    saturated, fed in random pieces."""
    rng = np.random.default_rng(seed)
    total = int(rng.integers(100_000, 3_000_000))
    gap = [3000, 700, 50_000, 1300, 641, 200_000][seed]
    starts, lens, n = [], [], int(rng.integers(0, 2000))
    while n < total - 1200:
        ln = 14 if rng.random() < 0.5 else 7
        starts.append(n)
        lens.append(ln)
        n += (1200 if ln == 14 else 640) + int(rng.integers(0, gap))
    starts, lens = np.array(starts, dtype=np.uint64), np.array(lens)
    keys = (starts << np.uint64(1)) | (lens == 14).astype(np.uint64)

    def model(state, n_have, scanned_upto, total_out):
        base, n_vis = state
        while True:
            call_at = base + 40980
            call_at += call_at & 1
            if call_at > total_out or call_at - 1200 > scanned_upto:
                return base, n_vis
            cutoff = call_at - 1200
            i = n_vis
            while i < n_have and int(starts[i]) < cutoff:
                i += 1
            pos = int(starts[i - 1]) + (1200 if lens[i - 1] == 14 else 640) if i > n_vis else base
            base, n_vis = max(cutoff, pos), i

    r = adsb.Resolver(True)
    state, n_vis, upto = (0, 0), 0, 0
    while upto < total:
        upto = min(total, upto + int(rng.integers(1, 400_000)))
        scanned = max(0, upto - 1200)
        have = int(np.searchsorted(starts, scanned))
        state = model(state, have, scanned, upto)
        n_vis = r.advance_keys(keys[:have], n_vis, scanned, upto)
        assert n_vis == state[1] and r.scan_end <= scanned
    assert r.advance_keys(keys[:have], n_vis, 0, 0, finish=True) == state[1]   # EOF adds no call (air.c:243-244)


def test_resolver_non_emulated_tail_scans_everything(adsb, oracle):
    cap = capture_basic(seed=32, seconds=0.03, n_frames=40, pad=0)
    ref = oracle.decode(cap.samples, want_amp=True)
    cands, _ = all_offset_candidates(ref.amp, False)
    total_out = ref.amp.size
    r = adsb.Resolver(False)
    r.feed(cands, total_out - 1200, total_out)
    r.finish()
    assert r.scan_end == total_out - 1200
    assert len(r.out) >= len(ref.frames)
    assert [(f.ts, f.data) for f in r.out][:len(ref.frames)] == [(f.ts, f.data) for f in ref.frames]


def test_plan_chunks_covers_and_aligns():
    from adsbdec_b200.chunking import plan_chunks
    for n_samples, k in ((2_000_000, 1), (2_000_000, 7), (40_000_000, 8), (9_000, 4), (2_400, 3)):
        ch = plan_chunks(n_samples, k)
        scan_end = max(0, n_samples // 2 - 1200)
        assert ch[0].own_begin == 0 and ch[-1].own_end == scan_end
        for a, b in zip(ch, ch[1:]):
            assert a.own_end == b.own_begin
        for c in ch:
            assert c.sample_begin % 8 == 0 and c.sample_end <= n_samples
            assert c.sample_end >= min(n_samples, 2 * (c.own_end + 1200))
            lo, hi = c.local_own
            assert (lo >= 8 or c.sample_begin == 0) and hi >= lo


def _chunk_candidates_cpu(oracle, samples, ch, df18):
    """Test double for Decoder.chunk_candidates: oracle envelope of the chunk (FIR phase = its
    absolute start), numpy per-offset logic, restricted to the owned range, LOCAL offsets."""
    seg = samples[ch.sample_begin:ch.sample_end]
    amp = oracle.decode(seg, want_amp=True, origin=ch.sample_begin).amp
    c, _ = all_offset_candidates(amp, df18)
    lo, hi = ch.local_own
    return c[(c["n"] >= lo) & (c["n"] < hi)]


@pytest.mark.parametrize("n_chunks", [1, 3, 8])
def test_chunked_merge_equals_whole(adsb, oracle, n_chunks):
    from adsbdec_b200.chunking import merge_candidates, plan_chunks, resolve_merged
    cap = capture_basic(seed=33, seconds=0.06, n_frames=90, pad=0)
    ref = oracle.decode(cap.samples, accept_df18=True)
    chunks = plan_chunks(cap.samples.size, n_chunks)
    parts = [(c, _chunk_candidates_cpu(oracle, cap.samples, c, True)) for c in chunks]
    merged = merge_candidates(parts)
    out = resolve_merged(merged, cap.samples.size)
    assert [(f.ts, f.pw, f.data) for f in out] == [(f.ts, f.pw, f.data) for f in ref.frames]


GLOO_WORKER = r'''
import os, sys, pickle
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import numpy as np
import torch.distributed as dist
from adsbdec_b200.chunking import plan_chunks, assign, merge_candidates, resolve_merged
from helpers import capture_basic
from test_host import _chunk_candidates_cpu
from oracle import pyoracle
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
cap = capture_basic(seed=34, seconds=0.05, n_frames=80, pad=0)
chunks = plan_chunks(cap.samples.size, 6)
mine = [(c, _chunk_candidates_cpu(pyoracle, cap.samples, c, True)) for c in assign(chunks, world, rank)]
gathered = [None] * world
dist.all_gather_object(gathered, mine)          # results only; the data path has no collective
if rank == 0:
    merged = merge_candidates([p for part in gathered for p in part])
    out = resolve_merged(merged, cap.samples.size)
    ref = pyoracle.decode(cap.samples, accept_df18=True)
    ok = [(f.ts, f.pw, f.data) for f in out] == [(f.ts, f.pw, f.data) for f in ref.frames]
    print("RESULT", ok, len(out), len(ref.frames))
dist.destroy_process_group()
'''


def test_two_rank_gloo_chunk_merge(tmp_path):
    """world_size 2 on CPU: ranks scan disjoint chunk sets, rank 0 merges by offset and resolves."""
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER.format(root=str(ROOT), tests=str(ROOT / "tests")))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert "RESULT True" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_pack12_layout_and_roundtrip(adsb):
    """The packed wire format of adsbdec_b200_push_packed12: 96 bits of three little-endian words, most
    significant first, are eight 12-bit samples (libairspy's unpack_samples)."""
    s = np.array([0xABC, 0x123, 0xFED, 0x001, 0x800, 0x7FF, 0x000, 0xFFF], dtype="<u2")
    w = adsb.pack12(s)
    assert w.dtype == np.dtype("<u4") and w.tolist() == [0xABC123FE, 0xD0018007, 0xFF000FFF]
    rng = np.random.default_rng(3)
    t = rng.integers(0, 4096, 8 * 1000).astype("<u2")
    assert np.array_equal(adsb.unpack12(adsb.pack12(t)), t)
    with pytest.raises(ValueError):
        adsb.pack12(np.full(8, 4096, dtype="<u2"))


def test_repair_table_against_brute_force(adsb, oracle):
    """EXTENSION: every pattern of one or two flipped bits in 112 has its own syndrome, and the library's hash
    table (host-built, the same one the kernels probe) returns exactly that pattern - checked with the
    modulator's bitwise CRC, which shares no code with it."""
    from adsbdec_b200.modulator import crc24

    def syn(bits):
        m = bytearray(14)
        for b in bits:
            m[b >> 3] ^= 0x80 >> (b & 7)
        return crc24(bytes(m[:11])) ^ int.from_bytes(m[11:], "big")
    one = [syn([b]) for b in range(112)]
    seen = {}
    for a in range(112):
        assert adsb.repair_lookup(one[a]) == (a << 8) | 0xFF
        seen[one[a]] = (a,)
        for b in range(a + 1, 112):
            s = one[a] ^ one[b]
            assert s not in seen, "two patterns share a syndrome"
            seen[s] = (a, b)
            assert adsb.repair_lookup(s) == (a << 8) | b
    assert len(seen) == 112 + 112 * 111 // 2
    assert syn([3, 77]) == one[3] ^ one[77]          # linearity, as the table assumes
    assert adsb.repair_lookup(0) == -1
    rng = np.random.default_rng(5)
    for s in rng.integers(1, 1 << 24, 2000):
        if int(s) not in seen:
            assert adsb.repair_lookup(int(s)) == -1
