"""Per-stage timing of the batched SIFT path (development aid; bench.py is the contract)."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import niftymatch_b200 as nm  # noqa: E402
from niftymatch_b200 import synth  # noqa: E402

w, h, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
peak = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
P = nm.SiftParams(w, h); P._peak_threshold = peak
frames = torch.from_numpy(synth.frame_batch(w, h, n)).cuda()
sb = nm.SiftBatch(P, n, 16384)
sb.enable_timing(True)
for it in range(4):
    sb.run(frames)
    torch.cuda.synchronize()
    ms = sb.stage_ms()
    print({k: round(v, 3) for k, v in ms.items()}, "frames/s=%.1f" % (n / ms["total"] * 1e3), "launches", sb.last_launches())
print("counts", sb.results()["counts"][:8].tolist())
