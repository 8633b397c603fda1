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
