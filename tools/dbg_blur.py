import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import niftymatch_b200 as nm
from niftymatch_b200 import sift as S
h, w, R = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
img = torch.rand((h, w), device="cuda") * 255
taps = torch.full((2 * R + 1,), 1.0 / (2 * R + 1), device="cuda")
out = S.blur(img, taps, R, buffer=torch.empty_like(img))
torch.cuda.synchronize()
print("ok", float(out.sum()), float(img.sum()))
