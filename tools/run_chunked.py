"""Multi-GPU time-chunked decode of ONE capture, checked against the oracle on rank 0.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_chunked.py"""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch, torch.distributed as dist
from adsbdec_b200.modulator import synth_capture
from adsbdec_b200.parallel import decode_chunked
from oracle import pyoracle

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if "RANK" in os.environ:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank() if dist.is_initialized() else 0
cap = synth_capture(2.0, 3000, snr_db=17, seed=77, df_mix=((17, .5), (11, .3), (18, .2)), pad=1 << 16)
t0 = time.perf_counter()
frames = decode_chunked(cap.samples, accept_df18=True)
dt = time.perf_counter() - t0
if rank == 0:
    ref = pyoracle.decode(cap.samples, accept_df18=True)
    same = [(f.ts, f.pw, f.data) for f in frames] == [(f.ts, f.pw, f.data) for f in ref.frames]
    world = dist.get_world_size() if dist.is_initialized() else 1
    print(f"world {world}: {len(frames)} frames, identical to the oracle: {same}, {dt*1e3:.1f} ms")
    assert same
if dist.is_initialized():
    dist.destroy_process_group()
