"""Layer-wise comparison of the fused kernel's taps (envelope, P bits, B bits) with the oracle."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import adsbdec_b200 as A
from adsbdec_b200.modulator import synth_capture
from oracle import pyoracle as po

origin = int(sys.argv[1]) if len(sys.argv) > 1 else 0
cap = synth_capture(0.25, 60, snr_db=20, seed=5, df_mix=((17, 0.6), (11, 0.3), (18, 0.1)), pad=1 << 16)
ref = po.decode(cap.samples, accept_df18=True, want_amp=True, origin=origin)
d = A.Decoder(accept_df18=True, start_sample=origin)
count = 400000
jb, amp, pb, qb, bb = d.debug_fused(cap.samples, count, origin)
print("j_begin", jb, "count", count)
a = ref.amp
lo = max(0, -jb)
seg = a[jb + lo: jb + count]
got = amp[lo:lo + seg.size]
bad = np.nonzero(seg.view(np.uint32) != got.view(np.uint32))[0]
print("amp mismatches:", bad.size, "of", seg.size, "first", bad[:10] + lo + jb)
if bad.size:
    i = bad[0]
    print(" ref", seg[i:i + 8], "\n got", got[i:i + 8])
c = (a[:-10] + a[10:]).astype(np.float32)
ci = c.astype(np.int64)
half = ci[:-5] > 2 * ci[5:]
qhalf = ci[5:] > 2 * ci[:-5]
bit = a[:-5] > a[5:]
nchk = count - 4000
for name, g, r in (("P", pb, half), ("Q", qb, qhalf), ("B", bb, bit)):
    rr = r[jb + lo: jb + nchk]
    gg = g[lo:nchk]
    bad = np.nonzero(rr != gg)[0]
    print(name, "bit mismatches:", bad.size, "of", rr.size, "first", bad[:10] + lo + jb, "density ref/got",
          rr.mean(), gg.mean())
for kernel in (0, 1):
    dd = A.decode(cap.samples, accept_df18=True, kernel=kernel, start_sample=origin)
    print("kernel", kernel, "frames", len(dd.frames), "oracle", len(ref.frames), dd.stats(), ref.tried, ref.ok)
