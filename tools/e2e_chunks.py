"""End-to-end (host -> host) frames/s of nm_sift_run_host for the pipeline chunk size in NM_HOST_CHUNK."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import niftymatch_b200 as nm
from niftymatch_b200 import synth
W, H, B = 1920, 1080, 64
base = synth.frame_batch(W, H, 8)
fr = np.stack([np.roll(base[i % 8], (3 * (i // 8), 5 * (i // 8)), axis=(0, 1)) for i in range(B)])
pin = torch.from_numpy(fr).pin_memory()
sb = nm.SiftBatch(nm.SiftParams(W, H), B, 16384)
out = sb.run_host(pin)
for _ in range(2): sb.run_host(pin, out=out)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(5): sb.run_host(pin, out=out)
dt = (time.perf_counter() - t) / 5
# raw copies for reference
dev = torch.empty((B, H, W), dtype=torch.float32, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(3): dev.copy_(pin, non_blocking=True)
torch.cuda.synchronize(); h2d = (time.perf_counter() - t) / 3
print(f"chunk={os.environ.get('NM_HOST_CHUNK', 'default')}: e2e {dt * 1e3:.2f} ms  {B / dt:.0f} frames/s   (H2D alone {h2d * 1e3:.2f} ms = {fr.nbytes / h2d / 1e9:.1f} GB/s)")
