"""One batched homography RANSAC call (64 pairs x 8192 correspondences x 1024 hypotheses), for ncu."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import niftymatch_b200 as nm  # noqa: E402

rng = np.random.default_rng(7)
sx = (rng.random((64, 8192)) * 1900).astype(np.float32)
sy = (rng.random((64, 8192)) * 1060).astype(np.float32)
dx = (1.01 * sx + 0.02 * sy + 5.0).astype(np.float32)
dy = (-0.02 * sx + 0.99 * sy - 3.0).astype(np.float32)
pts = [torch.from_numpy(a).cuda() for a in (sx, sy, dx, dy)]
for _ in range(2):
    H, st = nm.ransac_batch(nm.HOMOGRAPHY, *pts, None, 4.0, 1024, seed=1)
torch.cuda.synchronize()
print(st[:2].cpu().tolist())
