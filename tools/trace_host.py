import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import niftymatch_b200 as nm
from niftymatch_b200 import synth
W, H, B = 1920, 1080, 64
base = synth.frame_batch(W, H, 8)
fr = np.stack([np.roll(base[i % 8], (3 * (i // 8), 5 * (i // 8)), axis=(0, 1)) for i in range(B)])
pin = torch.from_numpy(fr).pin_memory()
sb = nm.SiftBatch(nm.SiftParams(W, H), B, 16384)
out = sb.run_host(pin)
for _ in range(3): sb.run_host(pin, out=out)
