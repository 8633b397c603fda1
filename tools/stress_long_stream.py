"""Stress: an 8.8 GB device-resident stream across the real 2^32 fidx wrap, fused vs generic kernels and the oracle tile property."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import adsbdec_b200 as A
from adsbdec_b200.modulator import synth_capture
DF_MIX_ALL = ((17, 0.5), (11, 0.25), (18, 0.15), (20, 0.05), (21, 0.05))
tile_cap = synth_capture(2.0, 5000, snr_db=16, seed=190, df_mix=DF_MIX_ALL, pad=0)
s = tile_cap.samples
s = s[:s.size - s.size % 28].copy(); s[-6000:] = 2048
reps = (2**32 + 60_000_000) // s.size + 1
big = torch.from_numpy(s.astype(np.int16)).cuda().repeat(reps)
out = []
for kernel in (0, 1):
    d = A.Decoder(accept_df18=True, kernel=kernel)
    d.push_device(big.data_ptr(), big.numel()); d.finish()
    out.append(d.frames_array); print(kernel, len(out[-1]), d.stats(), d.host_time()); d.close()
a, b = out
sa = {int(x) for x in a["n"]}; sb = {int(x) for x in b["n"]}
only_a = sorted(sa - sb); only_b = sorted(sb - sa)
W = 2**31
print("only fused:", [(n, n - W, n % 134217728) for n in only_a][:20])
print("only generic:", [(n, n - W, n % 134217728) for n in only_b][:20])
from oracle import pyoracle
ref = pyoracle.decode(np.concatenate([s, np.full(1 << 17, 2048, dtype="<u2")]), accept_df18=True)
tile_out = s.size // 2
refn = {f.n for f in ref.frames}
def chk(lst):
    return [(n, (n % tile_out) in refn) for n in lst]
print("tile_out", tile_out, "only fused in-oracle:", chk(only_a))
print("only generic in-oracle:", chk(only_b))
# does the fused output miss oracle frames in some tile?
for name, arr in (("fused", a), ("generic", b)):
    nn = arr["n"][arr["n"] < W - 2_000_000].astype(np.int64)
    t = nn // tile_out
    cnt = np.bincount(t)
    bad = np.nonzero(cnt != len(ref.frames))[0]
    print(name, "tiles with a different frame count:", bad[:20], cnt[bad][:20], "expected", len(ref.frames))
