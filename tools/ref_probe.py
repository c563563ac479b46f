import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from niftymatch_b200 import synth
from tests._util import load_reflib
ref = load_reflib()
w, h = int(sys.argv[1]), int(sys.argv[2])
img = synth.scene(w, h, synth.SEED_BASE)
t = time.time()
r = ref.sift_frame(img, peak=float(sys.argv[3]) if len(sys.argv) > 3 else 0.0)
print("ref n", r["n"], r["seg_counts"], time.time() - t, flush=True)
