"""Whole-step time of nm_sift_run without per-stage events (development aid).  usage: run_timing.py W H N [iters]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import niftymatch_b200 as nm  # noqa: E402
from niftymatch_b200 import synth  # noqa: E402

w, h, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 20
P = nm.SiftParams(w, h)
frames = torch.from_numpy(synth.frame_batch(w, h, n)).cuda()
sb = nm.SiftBatch(P, n, 16384)
for _ in range(5):
    sb.run(frames)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    sb.run(frames)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
c = sb.results()["counts"]
print("ms/step %.3f frames/s %.1f counts sum %d" % (ms, n / ms * 1e3, int(c.sum())))
