"""Small fused + generic decode against the oracle (for compute-sanitizer runs)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import adsbdec_b200 as A
from adsbdec_b200.modulator import synth_capture
from oracle import pyoracle
cap = synth_capture(0.3, 200, snr_db=18, seed=9, df_mix=((17, .5), (11, .3), (18, .2)), pad=1 << 16)
ref = pyoracle.decode(cap.samples, accept_df18=True)
for kernel in (0, 1):
    d = A.Decoder(accept_df18=True, kernel=kernel, batch_samples=1 << 21)
    d.push(cap.samples); d.finish()
    ok = [(f.n, f.ts, f.pw, f.data) for f in d.frames] == [(f.n, f.ts, f.pw, f.data) for f in ref.frames]
    print("kernel", kernel, "frames", len(d.frames), "== oracle:", ok, d.stats() == (ref.tried, ref.ok))
    d.close()

# the optional paths: repair extension, packed 12-bit input, live-source flush/poll
ref2 = pyoracle.decode(cap.samples, accept_df18=True, repair=2)
for kernel in (0, 1):
    d = A.Decoder(accept_df18=True, kernel=kernel, batch_samples=1 << 21, repair=2)
    n8 = cap.samples.size - cap.samples.size % 8
    d.push_packed12(A.pack12(cap.samples[:n8]))
    d.flush(); d.poll()
    d.push(cap.samples[n8:]); d.finish()
    ok = [(f.n, f.ts, f.pw, f.data) for f in d.frames] == [(f.n, f.ts, f.pw, f.data) for f in ref2.frames]
    print("kernel", kernel, "repair=2 packed12 frames", len(d.frames), "== oracle:", ok,
          [int(x) for x in d.frames_array["pad"][:, 0]] == ref2.fixed)
    d.close()

# the extensions of round 2 and the long-chain path of k_greedy
from adsbdec_b200.modulator import synth_fleet_capture, synth_iq_capture
key = lambda fr: [(f.n, f.ts, f.pw, f.data) for f in fr]
fleet = synth_fleet_capture(0.2, 250, 10, snr_db=22, seed=70, pad=1 << 16)
ref3 = pyoracle.decode(fleet.samples, accept_df18=True, address_parity=True)
for kernel in (0, 1):
    d = A.Decoder(accept_df18=True, kernel=kernel, batch_samples=1 << 20, address_parity=True)
    d.push(fleet.samples); d.finish()
    print("kernel", kernel, "address/parity frames", len(d.frames), "== oracle:", key(d.frames) == key(ref3.frames))
    d.close()
iq = synth_iq_capture(0.2, 200, snr_db=20, seed=80, df_mix=((17, .6), (11, .4)), pad=1 << 16)
ref4 = pyoracle.decode(iq.samples, accept_df18=True, iq=True)
d = A.Decoder(accept_df18=True, iq=True, batch_samples=1 << 20)
d.push(iq.samples); d.finish()
print("I/Q input frames", len(d.frames), "== oracle:", key(d.frames) == key(ref4.frames))
d.close()
sat = synth_capture(0.1, 1400, snr_db=22, seed=61, df_mix=((11, 1.0),), pad=1 << 16, poisson=False)
ref5 = pyoracle.decode(sat.samples, accept_df18=True)
d = A.Decoder(accept_df18=True)
d.push(sat.samples); d.finish()
print("saturated traffic (one greedy chain) frames", len(d.frames), "== oracle:", key(d.frames) == key(ref5.frames))
d.close()
