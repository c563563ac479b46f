#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native NiftyMatch hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): a batch of 64 synthetic 1920x1080 grayscale frames per
GPU, SIFT detect+describe with the reference's default parameters.  One step = one pass of the
hot path (nm_sift_run) over the batch.  N > 1 (torchrun, one rank per GPU): frames are sharded
per GPU, no data-path collective, weak scaling.  Prints ONE JSON line on rank 0.

  value      frames/s with the frames resident in HBM (CUDA events, max over ranks)
  e2e        frames/s through nm_sift_run_host: pinned host frames -> H2D -> run -> D2H of
             counts / descriptors / coordinates, every step
  roofline   the pyramid kernels (blur_tile_kernel<R>, the dominant HBM-bound kernel family)
             against the measured HBM copy bandwidth
  cpu_baseline  the CPU oracle (oracle/nm_oracle.c, a port: the reference has no CPU path)
             timed on this box's cores on a bounded sample
  match      auxiliary: brute-force k=2 matching of 100k x 100k descriptors (database sharded
             over the N ranks, NCCL all-gather of the per-shard top-2 records), Gpairs/s

--impl reference times the reference's own CUDA code (oracle/_ref/libnmref.so, built from
/root/reference by oracle/build_ref.sh) on the same frames; see DESIGN.md for why its
orientation step has to use the reference's kernel_orientations_naive.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, BATCH = 1920, 1080, 64
CAPACITY = 16384            # descriptor slots per frame (no frame truncates: ~6.4k keypoints)
N_OCT = 6
SUM_N = sum((W >> o) * (H >> o) for o in range(N_OCT))      # 2 764 020 pixels over the octaves
PYR_BYTES_PER_FRAME = 48 * SUM_N
EXT_BYTES_PER_FRAME = 24 * SUM_N
METRIC = "sift_frames_per_s_1080p"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for line in self.lines:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx = float(p[2])
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def measured_tensor_peak():
    """Dense bf16/fp16 TFLOP/s: the burst figure (the matcher is timed alone)."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["bf16_tflops"]), "measured burst (MEASURED_PEAKS.json bf16_tflops)"
    except Exception:
        return 1590.0, "fallback (B200_PROFILING.md)"


def traffic_from_profiles(kernel):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


def cpu_baseline(frames_np):
    """The CPU oracle on a bounded sample of the same frames, all host cores."""
    from tests._util import load_oracle
    import numpy as np
    orc = load_oracle()
    cores = os.cpu_count() or 1
    threads = min(cores, 64)
    n = min(BATCH, max(4, threads))
    sample = np.ascontiguousarray(frames_np[:n])
    cfg = np.array([0.0, -1, -1, CAPACITY, 1, 0], np.float32)
    counts = np.zeros(n, np.int32)
    t = time.perf_counter()
    orc.lib.orc_sift_batch(sample.ctypes.data_as(C.c_void_p), n, W, H, cfg.ctypes.data_as(C.c_void_p), threads,
                           counts.ctypes.data_as(C.c_void_p))
    dt = time.perf_counter() - t
    return {"value": n / dt, "unit": "frames/s", "cores": threads, "kind": "port",
            "sample": f"{n} of the {BATCH} 1080p frames, oracle/nm_oracle.c, {threads} OpenMP threads, {dt:.1f} s",
            "keypoints_per_frame": float(counts.mean())}


def run_reference_arm(args, rank, world):
    """The reference's own CUDA code on the same frames (rank 0 only)."""
    if rank != 0:
        return
    import numpy as np
    from niftymatch_b200 import synth
    base = {"impl": "reference", "metric": METRIC, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "batch of 1920x1080 synthetic frames, SIFT detect+describe (BASELINE.json configs[1])"}}
    n = 16
    frames = synth.frame_batch(W, H, n)
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libnmref.so")
    use_gpu = False
    try:
        import torch
        use_gpu = torch.cuda.is_available() and os.path.exists(ref_so)
    except Exception:
        pass
    if use_gpu:
        import torch
        lib = C.CDLL(ref_so)
        lib.nmref_sift_bench.restype = C.c_int
        dev = torch.from_numpy(frames).cuda()
        cfg = np.array([0.0, -1, -1, CAPACITY, 1, 1], np.float32)
        ms, items = C.c_float(), C.c_longlong()
        if args.warmup:
            lib.nmref_sift_bench(C.c_void_p(dev.data_ptr()), n, W, H, cfg.ctypes.data_as(C.c_void_p), args.warmup,
                                 C.byref(ms), C.byref(items))
        lib.nmref_sift_bench(C.c_void_p(dev.data_ptr()), n, W, H, cfg.ctypes.data_as(C.c_void_p), args.steps,
                             C.byref(ms), C.byref(items))
        per_step = ms.value / args.steps
        v = n / per_step * 1e3
        sample = (f"{n} of the {BATCH} frames per step on cuda:0 through the reference's own CUDA kernels "
                  "(oracle/_ref, unmodified sources; orientation by its kernel_orientations_naive because "
                  "detect_orientations deadlocks on sm_70+); one host thread")
        kind, cores = "reference", 1
        kp = items.value / (n * args.steps)
    else:
        cb = cpu_baseline(frames)
        v, per_step, sample, kind, cores, kp = cb["value"], n / cb["value"] * 1e3, cb["sample"], "port", cb["cores"], cb["keypoints_per_frame"]
    base.update({"value": v, "ms_per_step": per_step,
                 "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample,
                                  "keypoints_per_frame": kp},
                 "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    base["config"]["frames_per_step"] = n
    print(json.dumps(base), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-match", action="store_true", help="skip the auxiliary 100k x 100k matching measurement")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout must carry the one JSON line only: NCCL prints its version banner to stdout when the
        # communicator is created, so file descriptor 1 points at stderr until that has happened
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    import niftymatch_b200 as nm
    from niftymatch_b200 import synth
    from niftymatch_b200.dist import frame_range, shard_bounds

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- inputs: every rank owns its own 64 frames of the (64*N)-frame job ------------------
    lo, hi = frame_range(BATCH * world, world, rank)
    base_frames = synth.frame_batch(W, H, 8)                  # 8 distinct scenes (host render is slow)
    frames_np = np.empty((BATCH, H, W), np.float32)
    for i in range(BATCH):
        g = lo + i
        frames_np[i] = np.roll(base_frames[g % 8], (3 * (g // 8), 5 * (g // 8)), axis=(0, 1))
    frames_dev = torch.from_numpy(frames_np).cuda()
    frames_pinned = torch.from_numpy(frames_np).pin_memory()

    P = nm.SiftParams(W, H)
    sb = nm.SiftBatch(P, BATCH, CAPACITY)
    stream = torch.cuda.current_stream()

    # ---- device-resident timing -------------------------------------------------------------
    for _ in range(args.warmup):
        sb.run(frames_dev)
    barrier()
    sb.enable_timing(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    stage_acc = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        sb.run(frames_dev)
    e1.record(stream)
    if rank == 0:
        # the timed region lasts ~0.1 s, about one nvidia-smi period: keep the SAME load running (untimed, after the
        # closing event) until the sampler has seen it at least three times, so the clocks line is always "under load"
        t_lim = time.perf_counter() + 4.0
        while len(sampler.lines) < 3 and time.perf_counter() < t_lim:
            sb.run(frames_dev)
            torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    # stage events of the LAST step (all steps are identical work)
    stage = sb.stage_ms()
    sb.enable_timing(False)
    ms_step = max_over_ranks(ms_total / args.steps)
    value = BATCH * world / ms_step * 1e3
    launches_per_step = sb.last_launches()
    counts = sb.results()["counts"].cpu().numpy()

    # ---- end to end: pinned host frames -> H2D -> run -> D2H ---------------------------------
    out = sb.run_host(frames_pinned)                            # warm-up + allocates pinned outputs
    for _ in range(2):
        sb.run_host(frames_pinned, out=out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sb.run_host(frames_pinned, out=out)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    h2d = frames_np.nbytes * world
    d2h = sum_over_ranks(float(counts.sum()) * (128 + 2) * 4 + BATCH * 4)
    e2e = {"value": BATCH * world / e2e_ms * 1e3, "unit": "frames/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms}

    # ---- auxiliary: registration of 64 frame pairs (batched homography RANSAC), rank 0 only ----
    registration = None
    if rank == 0 and not args.no_match:
        try:
            rng = np.random.default_rng(7)
            n_pairs, n_pts, iters = 64, 8192, 1024
            sx = (rng.random((n_pairs, n_pts)) * 1900).astype(np.float32)
            sy = (rng.random((n_pairs, n_pts)) * 1060).astype(np.float32)
            dx = (1.01 * sx + 0.02 * sy + 5.0 + rng.normal(0, 0.3, sx.shape)).astype(np.float32)
            dy = (-0.02 * sx + 0.99 * sy - 3.0 + rng.normal(0, 0.3, sx.shape)).astype(np.float32)
            bad = rng.random(sx.shape) < 0.3
            dx[bad] = (rng.random(int(bad.sum())) * 1900).astype(np.float32)
            pts = [torch.from_numpy(a).cuda() for a in (sx, sy, dx, dy)]
            for _ in range(3):
                Hr, st_r = nm.ransac_batch(nm.HOMOGRAPHY, *pts, None, 4.0, iters, seed=1)
            torch.cuda.synchronize()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record(stream)
            for _ in range(5):
                Hr, st_r = nm.ransac_batch(nm.HOMOGRAPHY, *pts, None, 4.0, iters, seed=1)
            r1.record(stream)
            torch.cuda.synchronize()
            rms = r0.elapsed_time(r1) / 5
            registration = {"metric": "ransac_homography_pairs_per_s", "value": n_pairs / rms * 1e3, "unit": "pairs/s",
                            "ms_per_batch": rms, "config": {"workload": "64 frame pairs x 8192 correspondences (30 % outliers) x "
                                                            "1024 hypotheses, nm_ransac_batch_f32, no host round trip"},
                            "median_inliers": float(st_r[:, 1].float().median().item())}
            del pts
        except Exception as exc:                                 # auxiliary line: never fail the bench over it
            registration = {"error": repr(exc)}

    # ---- auxiliary: 100k x 100k matching, database sharded over the ranks --------------------
    match = None
    if not args.no_match:
        try:
            nq = ndb = 100000
            Bh = synth.descriptors(ndb, 2)
            Ah = synth.descriptors(nq, 1, planted_from=Bh)
            blo, bhi = shard_bounds(ndb, world, rank)
            A = torch.from_numpy(Ah).cuda()
            Bs = torch.from_numpy(np.ascontiguousarray(Bh[blo:bhi])).cuda()

            def match_step():
                if world > 1:
                    return nm.match_sharded(A, Bs, blo, 0.8)
                return nm.match(A, Bs, 0.8)

            msteps = max(2, min(args.steps, 5))
            for _ in range(2):
                m = match_step()
            barrier()
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record(stream)
            for _ in range(msteps):
                m = match_step()
            m1.record(stream)
            barrier()
            mms = max_over_ranks(m0.elapsed_time(m1) / msteps)
            tpeak, tsrc = measured_tensor_peak()
            gp = nq * ndb / mms / 1e6
            match = {"metric": "match_gpairs_per_s", "value": gp, "unit": "Gpairs/s",
                     "ms_per_step": mms, "scaling": "strong", "engine": nm.get_engine(),
                     "roofline": {"bound": "tensor", "achieved": gp * 256 / 1e3 / world, "peak": tpeak, "unit": "TFLOP/s",
                                  "frac": gp * 256 / 1e3 / world / tpeak, "peak_source": tsrc,
                                  "note": "256 flop per (query, database) pair = the 128-long -2a.b contraction "
                                          "(SURVEY.md 8d); per GPU; whole nm_match_f32 call (pack, tcgen05 scan, "
                                          "exact re-rank, fallback, merge), not the scan kernel alone",
                                  "traffic": None},
                     "config": {"workload": "100k x 100k 128-D fp32 descriptors, k=2 ratio test "
                                            "(BASELINE.json configs[3]); database rows sharded over the ranks, "
                                            "NCCL all-gather of per-shard top-2 records + merge",
                                "matched": int((m >= 0).sum().item())}}
        except Exception as exc:  # keep the headline line alive
            match = {"error": repr(exc)}

    # ---- CPU baseline (rank 0, N = 1 only) --------------------------------------------------
    cb = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cb = cpu_baseline(frames_np)
        except Exception as exc:
            cb = {"error": repr(exc)}

    if rank == 0:
        peak, peak_src = measured_peak()
        pyr_ms = stage["pyramid"]
        n_blur = 1 + 5 * N_OCT
        achieved = PYR_BYTES_PER_FRAME * BATCH / (pyr_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "blur_strip_kernel<R> (octaves 0-1) / blur_walk_kernel<R> / blur_tile_kernel<R> (31 launches per step: base + 5 levels x 6 octaves)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": PYR_BYTES_PER_FRAME * BATCH / n_blur,
                    "avg_launch_ms": pyr_ms / n_blur,
                    "traffic": traffic_from_profiles("blur_tile_kernel"),
                    "extrema_grad": {"achieved": EXT_BYTES_PER_FRAME * BATCH / (stage["extrema_grad"] * 1e-3) / 1e9,
                                     "unit": "GB/s", "algorithmic_bytes_per_launch_set": EXT_BYTES_PER_FRAME * BATCH}}
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "batch of 64 synthetic 1920x1080 frames per GPU, SIFT detect+describe, reference "
                                   "default parameters (BASELINE.json configs[1])",
                       "frames_per_gpu": BATCH, "capacity": CAPACITY, "keypoints_per_frame": float(counts.mean()),
                       "l2": "inputs (531 MB per step) larger than L2", "parallelism": f"frames sharded x{world}"},
            "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps * world),
            "launches_per_step": launches_per_step, "clocks": clocks, "stages_ms": stage,
            "roofline": roofline, "cpu_baseline": cb, "match": match, "registration": registration,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
