"""Achieved HBM bandwidth of the element-wise preprocessing kernels (development aid): algorithmic bytes
(read + written once) / CUDA-event time on 64 x 1080p, buffers far larger than L2."""
import ctypes as C
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import niftymatch_b200 as nm  # noqa: E402

lib = nm.load()
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6450.9
W, H, B = 1920, 1080, 64
n = W * H * B
bgra = torch.randint(0, 256, (B * H, W, 4), dtype=torch.uint8, device="cuda")
f32 = torch.rand((B * H, W), device="cuda") * 255
out = torch.empty((B * H, W), dtype=torch.float32, device="cuda")
u8 = torch.empty((B * H, W), dtype=torch.uint8, device="cuda")
yy, xx = torch.meshgrid(torch.arange(B * H, device="cuda", dtype=torch.float32), torch.arange(W, device="cuda", dtype=torch.float32), indexing="ij")
xx, yy = xx.contiguous(), yy.contiguous()
cam = torch.tensor([1400.0, 1400.0, 959.5, 539.5], device="cuda")
dist = torch.tensor([-0.2, 0.05, -0.003], device="cuda")
u, v = torch.empty_like(xx), torch.empty_like(yy)
p = lambda t: C.c_void_p(t.data_ptr())
cases = [
    ("nm_grayscale_bgra_f32", 8 * n, lambda: lib.nm_grayscale_bgra_f32(p(bgra), p(out), W, B * H, None)),
    ("nm_cast_f32_u8", 5 * n, lambda: lib.nm_cast_f32_u8(p(f32), W, B * H, p(u8), 0, None)),
    ("nm_undistort_map_f32", 16 * n, lambda: lib.nm_undistort_map_f32(p(xx), p(yy), W, B * H, p(cam), p(dist), p(u), p(v), None)),
    ("nm_bgra_extract_channel_f32", 8 * n, lambda: lib.nm_bgra_extract_channel_f32(p(bgra), p(out), W, B * H, 1, None)),
]
for name, nbytes, fn in cases:
    for _ in range(3):
        assert fn() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gbs = nbytes / ms / 1e6
    print(f"{name}: {ms:.3f} ms per 64 x 1080p = {gbs:.0f} GB/s algorithmic = {gbs / peak:.2f} of the measured HBM peak ({peak:.0f} GB/s)")
