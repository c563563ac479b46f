#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native NiftyMatch hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): a batch of 64 synthetic 1920x1080 grayscale frames per
GPU (frame g of the job rendered from seed 0x5EED0000 + g), SIFT detect+describe
with the reference's default parameters.  One step = one pass of the hot path (nm_sift_run) over
the batch.  N > 1 (torchrun, one rank per GPU): frames are sharded per GPU, no data-path
collective, weak scaling.  Prints ONE JSON line on rank 0.

  value        frames/s with the frames resident in HBM (CUDA events, max over ranks)
  e2e          frames/s through nm_sift_run_host: pinned host frames -> H2D -> run -> D2H of
               counts / descriptors / coordinates, every step; copy_ceiling_ms = the same bytes
               moved by bare cudaMemcpyAsync calls (H2D and D2H concurrently, all ranks at once)
  roofline     the dominant kernel of the step (largest measured stage kernel) against the measured
               HBM copy bandwidth; `stages` = every stage with its algorithmic bytes and fraction;
               the blur family (pyramid) as a sub-entry
  cpu_baseline the CPU oracle (oracle/nm_oracle.c, a port: the reference has no CPU path) timed on
               this box's cores on a bounded sample
  config1      BASELINE.json configs[0]: one 640x480 frame pair, detect+describe + k=2 match
  config3      BASELINE.json configs[2]: 8 x 3840x2160, 6 octaves, pyramid + extrema stages
  match        BASELINE.json configs[3]: 100k x 100k k=2 matching, database sharded over the N
               ranks (NCCL all-gather of per-shard top-2 records), Gpairs/s; `sizes` = 8k / 20k /
               40k square problems, the ones the reference's 32-bit indexing can address
  stream       BASELINE.json configs[4]: mosaicking stream, SIFT + consecutive-frame matching +
               registration, contiguous frame ranges per GPU with one overlap frame

--impl reference times the reference's own CUDA code (oracle/_ref/libnmref.so, built from
/root/reference by oracle/build_ref.sh) on the same frames and its compute_sift_matches at the
8k / 20k / 40k sizes; see DESIGN.md for the two places where it cannot run unmodified.
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
CAPACITY = 16384            # descriptor slots per frame (no frame truncates: ~6.3k keypoints)
N_OCT = 6
SUM_N = sum((W >> o) * (H >> o) for o in range(N_OCT))      # 2 764 020 pixels over the octaves
PYR_BYTES_PER_FRAME = 48 * SUM_N
EXT_BYTES_PER_FRAME = 24 * SUM_N
METRIC = "sift_frames_per_s_1080p"
MATCH_SIZES = (8000, 20000, 40000)                            # N x N problems the reference can index (N^2 < 2^31)
CONFIG = {"workload": "batch of 64 synthetic 1920x1080 frames per GPU (seed 0x5EED0000 + frame), SIFT detect+describe, "
                      "reference default parameters (BASELINE.json configs[1])",
          "frames_per_gpu": BATCH, "capacity": CAPACITY, "l2": "inputs (531 MB per step) larger than L2"}


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
    """dram bytes per launch of a kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


# ---- synthetic frames: one scene per seed, rendered on the host cores before CUDA is touched ----------------
def _render(job):
    from niftymatch_b200 import synth
    w, h, seed, shift = job
    return synth.scene(w, h, seed, shift=shift)


def render_frames(jobs, world=1):
    """jobs: list of (w, h, seed, shift).  Forked workers (numpy only; must run before the CUDA context exists)."""
    import multiprocessing as mp
    import numpy as np
    workers = max(1, min(16, (os.cpu_count() or 2) // max(world, 1), len(jobs)))
    if workers == 1:
        return np.stack([_render(j) for j in jobs])
    with mp.get_context("fork").Pool(workers) as pool:
        return np.stack(pool.map(_render, jobs, chunksize=1))


def frame_jobs(lo, n, w=W, h=H):
    from niftymatch_b200 import synth
    return [(w, h, synth.SEED_BASE + lo + i, (0.0, 0.0)) for i in range(n)]


def stream_jobs(lo, n, w=W, h=H, scene_len=16):
    """Mosaicking stream (configs[4]): frame t shows scene t // scene_len translated by a sub-pixel drift that grows
    with t % scene_len (<= 8 px), fresh sensor noise per frame, so that consecutive frames have true matches."""
    from niftymatch_b200 import synth
    return [(w, h, synth.SEED_BASE + 0x10000 + (lo + i) // scene_len,
             (0.4375 * ((lo + i) % scene_len), 0.3125 * ((lo + i) % scene_len))) for i in range(n)]


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
            "dtype": "f32", "data": "synthetic", "config": dict(CONFIG, parallelism=f"frames sharded x{world}")}
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libnmref.so")
    frames = render_frames(frame_jobs(0, BATCH))
    use_gpu = False
    try:
        import torch
        use_gpu = torch.cuda.is_available() and os.path.exists(ref_so)
    except Exception:
        pass
    match_sizes = None
    if use_gpu:
        import torch
        lib = C.CDLL(ref_so)
        lib.nmref_sift_bench.restype = C.c_int
        dev = torch.from_numpy(frames).cuda()
        # orientation mode 3 = the reference's public kernel with its two divergent barriers hoisted (it deadlocks on
        # sm_70+ as shipped; oracle/build_ref.sh); clear_grad = 1 as in the parity runs
        cfg = np.array([0.0, -1, -1, CAPACITY, 1, 3], np.float32)
        ms, items = C.c_float(), C.c_longlong()
        n = BATCH
        if args.warmup:
            lib.nmref_sift_bench(C.c_void_p(dev.data_ptr()), n, W, H, cfg.ctypes.data_as(C.c_void_p), args.warmup,
                                 C.byref(ms), C.byref(items))
        rc = lib.nmref_sift_bench(C.c_void_p(dev.data_ptr()), n, W, H, cfg.ctypes.data_as(C.c_void_p), args.steps,
                                  C.byref(ms), C.byref(items))
        assert rc == 0, rc
        per_step = ms.value / args.steps
        v = n / per_step * 1e3
        sample = (f"all {n} frames of a step on cuda:0 through the reference's own CUDA kernels (oracle/_ref, its sources "
                  "compiled where they lie; orientation = its public kernel with the two divergent __syncthreads() "
                  "hoisted, because detect_orientations deadlocks on sm_70+ as shipped); one host thread")
        kind, cores = "reference", 1
        kp = items.value / (n * args.steps)
        # for the record: the same run with the reference's OTHER orientation kernel (kernel_orientations_naive, one thread
        # per keypoint, no window clamp) -- round 1 timed that one, because the public kernel cannot run unpatched
        alt = None
        try:
            cfg1 = np.array([0.0, -1, -1, CAPACITY, 1, 1], np.float32)
            ms1, it1 = C.c_float(), C.c_longlong()
            if lib.nmref_sift_bench(C.c_void_p(dev.data_ptr()), n, W, H, cfg1.ctypes.data_as(C.c_void_p), 2, C.byref(ms1), C.byref(it1)) == 0:
                alt = {"frames_per_s": n / (ms1.value / 2) * 1e3, "orientation_kernel": "kernel_orientations_naive (orientation.cu:132-216)"}
        except Exception as exc:
            alt = {"error": repr(exc)}
        del dev
        # the matcher half of the metric at the sizes the reference's 32-bit indexing can address
        try:
            lib.nmref_match_bench.restype = C.c_int
            match_sizes = []
            for nn in MATCH_SIZES:
                Bh = synth.descriptors(nn, 2)
                Ah = synth.descriptors(nn, 1, planted_from=Bh)
                A, B = torch.from_numpy(Ah).cuda(), torch.from_numpy(Bh).cuda()
                mm = C.c_float()
                iters = 3
                rc = lib.nmref_match_bench(C.c_void_p(A.data_ptr()), nn, C.c_void_p(B.data_ptr()), nn, C.c_float(0.8), iters,
                                           C.byref(mm))
                assert rc == 0, rc
                match_sizes.append({"n": nn, "ms": mm.value / iters, "gpairs_per_s": nn * nn / (mm.value / iters) / 1e6})
                del A, B
                torch.cuda.empty_cache()
        except Exception as exc:
            match_sizes = {"error": repr(exc)}
    else:
        cb = cpu_baseline(frames)
        n = BATCH
        v, per_step, sample, kind, cores, kp = cb["value"], n / cb["value"] * 1e3, cb["sample"], "port", cb["cores"], cb["keypoints_per_frame"]
    if use_gpu:
        base["with_naive_orientation_kernel"] = alt
    base.update({"value": v, "ms_per_step": per_step, "keypoints_per_frame": kp,
                 "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
                 "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "match": {"metric": "match_gpairs_per_s", "unit": "Gpairs/s", "sizes": match_sizes,
                           "note": "compute_sift_matches of the reference (transpose + brute-force distance matrix + "
                                   "row scan, match.cu / siftfunctions.cu:15-40), CHUNK forced to one value; 100k x 100k "
                                   "is out of its reach: int indices (match.cu:89, siftfunctions.cu:28) and 80 GB of matrices"}})
    print(json.dumps(base), flush=True)


def window_sample_bytes(kpts, counts, meta_oct_dims):
    """Gradient samples (8 B each) the orientation / descriptor windows of the emitted keypoints read, from the
    keypoint payloads: the algorithmic bytes of the two gather kernels (descriptor.cu:54-65, orientation.cu:26-46)."""
    import numpy as np
    ori = desc = 0
    n_kp = 0
    for f in range(kpts.shape[0]):
        k = kpts[f, : counts[f]]
        k = k[k[:, 3] >= 0]
        if not len(k):
            continue
        n_kp += len(k)
        # octave of a keypoint from its scale: sigma = sigma_0 * 2^((l + ds)/3) * 2^o, l + ds in (-1, 3)
        o = np.clip(np.floor(np.log2(k[:, 2] / 2.0159) - (-1.0 / 3)).astype(int), 0, len(meta_oct_dims) - 1)
        xper = 2.0 ** o
        x, y, s = k[:, 0] / xper, k[:, 1] / xper, k[:, 2] / xper
        xi, yi = (x + 0.5).astype(int), (y + 0.5).astype(int)
        ow = np.array([meta_oct_dims[i][0] for i in o]); oh = np.array([meta_oct_dims[i][1] for i in o])
        wo = np.minimum(10, np.maximum(np.floor(4.5 * s), 1)).astype(int)
        ori += int(((np.minimum(wo, ow - 1 - xi) - np.maximum(-wo, -xi) + 1).clip(0) *
                    (np.minimum(wo, oh - 1 - yi) - np.maximum(-wo, -yi) + 1).clip(0)).sum())
        wd = np.floor(np.sqrt(2.0) * (3 * s + 1e-7) * 2.5 + 0.5).astype(int)
        xmin, xmax = np.maximum(-wd, -xi), np.minimum(wd, ow - 1 - xi)
        ymin, ymax = np.maximum(-wd, -yi), np.minimum(wd, oh - 1 - yi)
        ch = np.ceil((np.maximum(xmax - xmin, ymax - ymin) + 1.0) / 16).astype(int)
        for c in range(int(ch.max()) if len(ch) else 0):
            live = ch > c
            nx = (np.minimum(xmin + 16 * c + 15, xmax) - (xmin + 16 * c) + 1).clip(0)
            ny = (np.minimum(ymin + 16 * c + 15, ymax) - (ymin + 16 * c) + 1).clip(0)
            desc += int((nx * ny * live).sum())
    return ori * 8 + n_kp * 36, desc * 8 + n_kp * (512 + 8 + 28), n_kp


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-match", action="store_true", help="skip the auxiliary matching / registration / stream legs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[0] / configs[2] legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import numpy as np
    from niftymatch_b200.dist import frame_range, shard_bounds
    from niftymatch_b200 import synth

    # ---- inputs: rendered before the CUDA context exists (forked workers) ------------------------------------
    lo, hi = frame_range(BATCH * world, world, rank)
    t_r = time.perf_counter()
    frames_np = render_frames(frame_jobs(lo, BATCH), world)          # frame g of the job: seed 0x5EED0000 + g
    extra_np = {}
    if rank == 0 and world == 1 and not args.no_extra:
        extra_np["c1"] = render_frames([(640, 480, synth.SEED_BASE, (0.0, 0.0)), (640, 480, synth.SEED_BASE, (2.5, 1.25))])
        extra_np["c3"] = render_frames(frame_jobs(0, 8, 3840, 2160))
    STREAM_FRAMES = env_int("NM_BENCH_STREAM_FRAMES", 2048)             # whole stream, split over the ranks
    slo, shi = frame_range(STREAM_FRAMES, world, rank, overlap=1)
    # the stream's frames are rendered at chunk granularity on the device side of the leg (host render of 2048 frames
    # would take minutes): 32 distinct host-rendered frames, drifted per stream position (see stream_leg)
    stream_np = None if args.no_match else render_frames(stream_jobs(0, 32), world)
    render_s = time.perf_counter() - t_r

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
    res = sb.results()
    counts = res["counts"].cpu().numpy()
    kpts_np = res["kpts"].cpu().numpy() if rank == 0 else None

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
    d2h_rank = float(counts.sum()) * (128 + 2) * 4 + BATCH * 4
    d2h = sum_over_ranks(d2h_rank)
    # the ceiling of that path: the same bytes by bare cudaMemcpyAsync (one per buffer), H2D and D2H on two streams at
    # once, every rank at the same time -- what the host's memory system and the PCIe links give, no kernels
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    dn_elems = int(d2h_rank // 4)
    dn_src = torch.empty(dn_elems, dtype=torch.float32, device="cuda")
    dn_dst = torch.empty(dn_elems, dtype=torch.float32).pin_memory()
    up_dst = torch.empty_like(frames_dev)

    def copies(pieces):
        # `pieces` cudaMemcpyAsync calls per direction (1 = one call per buffer; more = the granularity a pipeline uses)
        with torch.cuda.stream(s_up):
            for a, b in zip(up_dst.view(-1).chunk(pieces), frames_pinned.view(-1).chunk(pieces)):
                a.copy_(b, non_blocking=True)
        with torch.cuda.stream(s_dn):
            for a, b in zip(dn_dst.chunk(pieces), dn_src.chunk(pieces)):
                a.copy_(b, non_blocking=True)

    ceil_ms, ceil_pieces = None, 1
    for pieces in (1, 16):
        for _ in range(2):
            copies(pieces)
        s_up.synchronize(); s_dn.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            copies(pieces)
        s_up.synchronize(); s_dn.synchronize()
        barrier()
        ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
        if ceil_ms is None or ms < ceil_ms:
            ceil_ms, ceil_pieces = ms, pieces
    del dn_src, dn_dst, up_dst
    e2e = {"value": BATCH * world / e2e_ms * 1e3, "unit": "frames/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "copy_ceiling_ms": ceil_ms,
           "frac_of_copy_ceiling": ceil_ms / e2e_ms,
           "copy_ceiling_note": "bare cudaMemcpyAsync of the step's H2D and D2H bytes from / to pinned memory, two streams, "
                                f"all ranks at once, the faster of 1 and 16 calls per buffer ({ceil_pieces}); aggregate GB/s = " +
                                f"{h2d / ceil_ms / 1e6:.1f} up + {d2h / ceil_ms / 1e6:.1f} down"}

    peak, peak_src = measured_peak()
    extra = {}
    # ---- configs[0]: one 640x480 pair, detect + describe + match; configs[2]: 8 x 4K, pyramid + extrema ----------
    if "c1" in extra_np:
        try:
            extra["config1"] = config1_leg(nm, torch, extra_np["c1"], peak, args)
        except Exception as exc:
            extra["config1"] = {"error": repr(exc)}
        try:
            extra["config3"] = config3_leg(nm, torch, extra_np["c3"], peak, args)
        except Exception as exc:
            extra["config3"] = {"error": repr(exc)}

    # ---- auxiliary: registration of 64 frame pairs (batched homography RANSAC), rank 0 only ----
    registration = None
    if rank == 0 and world == 1 and not args.no_match:
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

    # ---- configs[3]: 100k x 100k matching, database sharded over the ranks --------------------
    match = None
    if not args.no_match:
        try:
            match = match_leg(nm, torch, dist, synth, shard_bounds, rank, world, args, barrier, max_over_ranks, stream)
        except Exception as exc:  # keep the headline line alive
            match = {"error": repr(exc)}

    # ---- configs[4]: the mosaicking stream ------------------------------------------------------
    stream_line = None
    if not args.no_match:
        try:
            stream_line = stream_leg(nm, torch, dist, sb, stream_np, STREAM_FRAMES, slo, shi, rank, world, barrier, max_over_ranks)
        except Exception as exc:
            stream_line = {"error": repr(exc)}

    # ---- CPU baseline (rank 0, N = 1 only) --------------------------------------------------
    cb = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cb = cpu_baseline(frames_np)
        except Exception as exc:
            cb = {"error": repr(exc)}

    if rank == 0:
        dims = [(W >> o, H >> o) for o in range(N_OCT)]
        ori_b, desc_b, n_kp = window_sample_bytes(kpts_np, counts, dims)
        n_blur = 1 + 5 * N_OCT

        def st(name, kernel, ms, nbytes, launches, note):
            ach = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            return {"stage": name, "kernel": kernel, "ms": ms, "launches": launches, "bound": "hbm",
                    "algorithmic_bytes": int(nbytes), "achieved": ach, "unit": "GB/s", "frac": ach / peak, "note": note}

        stages = [
            st("pyramid", "blur_stream_kernel<R> (octaves 0-1) / blur_walk_kernel<R> / blur_tile_kernel<R>", stage["pyramid"],
               PYR_BYTES_PER_FRAME * BATCH, n_blur, "48 B per pixel and octave: 6 levels written once, one level-sized read per blur (SURVEY.md 8d)"),
            st("extrema", "extrema_kernel + refine_list_kernel", stage["extrema"], EXT_BYTES_PER_FRAME * BATCH, 2 * N_OCT,
               "24 B per pixel and octave: each of the 6 levels read once; DoG never materialised"),
            st("compaction", "rank_kernel, plan_kernel, emit_kernel, kprefine_kernel", stage["compaction"],
               BATCH * 3 * SUM_N / 8 * 3 + n_kp * 40, 3 + N_OCT, "bitmaps read twice + word prefixes written; latency bound, no roofline target"),
            st("gradient", "gradmap_kernel", stage["gradient"], 36 * SUM_N * BATCH, N_OCT,
               "design overhead, not credited by SURVEY.md 8d: 3 levels read (12 B) + 3 float2 maps written (24 B) per pixel if every "
               "block were computed; only blocks that keypoint windows read are"),
            st("orientation", "orient_kernel", stage["orientation"], ori_b, 1, "gather: 8 B per window sample of every keypoint + payloads; keypoint-count bound"),
            st("descriptor", "describe_fast_kernel<16, 4>", stage["descriptor"], desc_b, 1, "gather: 8 B per sample of the diagonal 16x16 chunks + 512 B descriptor per keypoint"),
        ]
        single = [s for s in stages if s["stage"] in ("extrema", "gradient", "orientation", "descriptor")]
        dom = max(single, key=lambda s: s["ms"])
        pyr = stages[0]
        dom_is_pyr = pyr["ms"] / 5.0 > dom["ms"]            # the largest blur instantiation (R = 7: 3 of the 11 large launches)
        top = pyr if dom_is_pyr else dom
        roofline = {"bound": "hbm", "kernel": top["kernel"], "stage": top["stage"], "achieved": top["achieved"], "peak": peak,
                    "unit": "GB/s", "frac": top["frac"], "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": top["algorithmic_bytes"] / top["launches"],
                    "avg_launch_ms": top["ms"] / top["launches"],
                    "traffic": traffic_from_profiles(top["stage"]),
                    "ncu": (traffic_from_profiles("ncu_limits") or {}).get(top["stage"]),
                    "note": "dominant kernel = the stage kernel with the largest CUDA-event time in this run (the pyramid is 31 "
                            "launches of five blur instantiations, none of them larger); `ncu` = what the committed ncu page of "
                            "this kernel shows as its limiter (profiles/r02_kernels.md): the gather kernels are bound by "
                            "instruction issue, so their HBM fraction stays low by construction; " + top["note"],
                    "blur_family": {k: pyr[k] for k in ("kernel", "ms", "launches", "algorithmic_bytes", "achieved", "frac")},
                    "dense_stages": {"ms": stage["pyramid"] + stage["extrema"],
                                     "achieved": 72 * SUM_N * BATCH / ((stage["pyramid"] + stage["extrema"]) * 1e-3) / 1e9,
                                     "frac": 72 * SUM_N * BATCH / ((stage["pyramid"] + stage["extrema"]) * 1e-3) / 1e9 / peak,
                                     "note": "pyramid + DoG/extrema against 72 B per pixel and octave (SURVEY.md 8d)"}}
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(CONFIG, parallelism=f"frames sharded x{world}"),
            "keypoints_per_frame": float(counts.mean()),
            "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps * world),
            "launches_per_step": launches_per_step, "clocks": clocks, "stages_ms": stage, "stages": stages,
            "roofline": roofline, "cpu_baseline": cb, "match": match, "registration": registration,
            "stream": stream_line, "host_render_s": render_s,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _time_events(torch, fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def config1_leg(nm, torch, pair_np, peak, args):
    """BASELINE.json configs[0]: a 640x480 frame and its shifted pair, SIFT detect+describe on both and the 2-frame k=2
    ratio-test match; the CPU oracle single-threaded beside it (the "CPU baseline reference run")."""
    import numpy as np
    from tests._util import load_oracle
    w, h = 640, 480
    P = nm.SiftParams(w, h)
    sbp = nm.SiftBatch(P, 2, 32768)
    dev = torch.from_numpy(pair_np).cuda()
    pinned = torch.from_numpy(pair_np).pin_memory()
    sbp.run(dev)
    torch.cuda.synchronize()
    r = sbp.results()
    n0, n1 = int(r["counts"][0].item()), int(r["counts"][1].item())

    def step():
        sbp.run(dev)
        return nm.match(r["desc"][0, :n0], r["desc"][1, :n1], 0.8)

    ms = _time_events(torch, step, max(args.steps, 10), 3)
    sbp.enable_timing(True)
    sbp.run(dev)
    torch.cuda.synchronize()
    stg = sbp.stage_ms()
    sbp.enable_timing(False)
    m = step()
    torch.cuda.synchronize()
    out = sbp.run_host(pinned)
    t0 = time.perf_counter()
    reps = max(args.steps, 10)
    for _ in range(reps):
        out = sbp.run_host(pinned, out=out)
        dd = torch.from_numpy(out["desc"].numpy()[:, : max(n0, n1)]).cuda(non_blocking=True)
        mm = nm.match(dd[0, :n0], dd[1, :n1], 0.8).cpu()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / reps
    sum_n = sum((w >> o) * (h >> o) for o in range(P._num_octaves))
    dense_ms = stg["pyramid"] + stg["extrema"]
    ach = 72 * sum_n * 2 / (dense_ms * 1e-3) / 1e9
    # CPU oracle, one thread
    orc = load_oracle()
    t0 = time.perf_counter()
    c0 = orc.sift_frame(pair_np[0], want_levels=False, capacity=32768)
    c1 = orc.sift_frame(pair_np[1], want_levels=False, capacity=32768)
    mo = orc.match(c0["desc"], c1["desc"], 0.8)
    cpu_s = time.perf_counter() - t0
    sbp.close()
    return {"workload": "one 640x480 frame + its shifted pair (seed 0x5EED0000, shift 2.5 / 1.25 px): detect+describe both, "
                        "k=2 ratio-test match (BASELINE.json configs[0])",
            "ms_per_pair": ms, "pairs_per_s": 1e3 / ms, "frames_per_s": 2e3 / ms, "e2e_ms_per_pair": e2e_ms,
            "keypoints": [n0, n1], "matched": int((m >= 0).sum().item()), "stages_ms": stg,
            "roofline": {"bound": "hbm", "stage": "pyramid + extrema (dense stages)", "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "note": "2 frames = 59 MB of algorithmic traffic: a launch-latency-sized problem "
                                                     "(28 launches, ~0.1 ms), not a bandwidth one"},
            "cpu_oracle": {"s_per_pair": cpu_s, "cores": 1, "kind": "port", "keypoints": [int(c0["n"]), int(c1["n"])],
                           "matched": int((mo >= 0).sum()), "same_match_indices": bool(np.array_equal(mo, m.cpu().numpy()))},
            "speedup_vs_cpu_oracle": cpu_s * 1e3 / ms}


def config3_leg(nm, torch, frames4k_np, peak, args):
    """BASELINE.json configs[2]: 8 x 3840x2160, 6 octaves forced (the default would be 7); the bandwidth-bound stages
    (pyramid + DoG/extrema) are what is reported, the rest of the step runs but is not counted."""
    w, h, n = 3840, 2160, frames4k_np.shape[0]
    P = nm.SiftParams(w, h)
    P._num_octaves = 6
    sb4 = nm.SiftBatch(P, n, 65536)
    dev = torch.from_numpy(frames4k_np).cuda()
    for _ in range(3):
        sb4.run(dev)
    torch.cuda.synchronize()
    sb4.enable_timing(True)
    acc = {"pyramid": 0.0, "extrema": 0.0, "total": 0.0}
    reps = max(args.steps, 5)
    for _ in range(reps):
        sb4.run(dev)
        torch.cuda.synchronize()
        s = sb4.stage_ms()
        for k in acc:
            acc[k] += s[k] / reps
    counts = sb4.results()["counts"][:n].cpu().numpy()
    sb4.close()
    sum_n = sum((w >> o) * (h >> o) for o in range(6))
    pyr = 48 * sum_n * n / (acc["pyramid"] * 1e-3) / 1e9
    ext = 24 * sum_n * n / (acc["extrema"] * 1e-3) / 1e9
    both = 72 * sum_n * n / ((acc["pyramid"] + acc["extrema"]) * 1e-3) / 1e9
    return {"workload": "8 synthetic 3840x2160 frames, 6-octave pyramid + DoG/extrema stages (BASELINE.json configs[2])",
            "frames_per_s_dense_stages": n / (acc["pyramid"] + acc["extrema"]) * 1e3, "frames_per_s_whole_step": n / acc["total"] * 1e3,
            "stages_ms": acc, "keypoints_per_frame": float(counts.mean()),
            "roofline": {"bound": "hbm", "peak": peak, "unit": "GB/s",
                         "pyramid": {"achieved": pyr, "frac": pyr / peak, "algorithmic_bytes": 48 * sum_n * n},
                         "extrema": {"achieved": ext, "frac": ext / peak, "algorithmic_bytes": 24 * sum_n * n},
                         "dense_stages": {"achieved": both, "frac": both / peak, "algorithmic_bytes": 72 * sum_n * n}}}


def match_leg(nm, torch, dist, synth, shard_bounds, rank, world, args, barrier, max_over_ranks, stream):
    import numpy as np
    nq = ndb = 100000
    Bh = synth.descriptors(ndb, 2)
    Ah = synth.descriptors(nq, 1, planted_from=Bh)
    # world = Q x D: Q query groups, D database shards (rank r scans query block r // D against shard r % D); Q = 1 is
    # pure database sharding.  The per-row costs of a scan shrink with Q, the scan itself with D.
    qg = env_int("NM_BENCH_QGROUPS", {8: 4, 4: 4, 2: 2}.get(world, 1))   # measured best grids: 2x1, 4x1, 4x2
    if world % qg:
        qg = 1
    n_db_shards = world // qg
    blo, bhi = shard_bounds(ndb, n_db_shards, rank % n_db_shards)
    A = torch.from_numpy(Ah).cuda()
    Bs = torch.from_numpy(np.ascontiguousarray(Bh[blo:bhi])).cuda()

    mg = None
    if world > 1:
        # the C-ABI multi-GPU entry (include/nm_b200_mgpu.h): its own NCCL communicator, created from a unique id that
        # rank 0 makes and torch.distributed carries to the other ranks
        from niftymatch_b200 import mgpu
        uid_t = torch.tensor(list(mgpu.unique_id() if rank == 0 else bytes(mgpu.ID_BYTES)), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid_t, 0)
        mg = mgpu.MultiGpu(rank=rank, world=world, uid=bytes(uid_t.cpu().tolist()))
        mg.set_query_groups(qg)
        io = torch.full((nq,), -1, dtype=torch.int32, device="cuda")

    def match_step():
        if mg is not None:
            return mg.match([A], [Bs], [blo], 0.8, match_io=[io], streams=[torch.cuda.current_stream()])[0]
        return nm.match(A, Bs, 0.8)

    msteps = max(3, min(args.steps, 10))
    for _ in range(3):
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
    mh = m.cpu().numpy()
    index_hash = int(np.bitwise_xor.reduce((mh.astype(np.int64) + 2) * (np.arange(nq, dtype=np.int64) * 2654435761 % (1 << 31))))
    out = {"metric": "match_gpairs_per_s", "value": gp, "unit": "Gpairs/s",
           "ms_per_step": mms, "scaling": "strong", "engine": nm.get_engine(),
           "roofline": {"bound": "tensor", "achieved": gp * 256 / 1e3 / world, "peak": tpeak, "unit": "TFLOP/s",
                        "frac": gp * 256 / 1e3 / world / tpeak, "peak_source": tsrc,
                        "note": "256 flop per (query, database) pair = the 128-long -2a.b contraction "
                                "(SURVEY.md 8d); per GPU; whole call (pack, tcgen05 scan, exact re-rank, fallback, "
                                "all-gather, merge), not the scan kernel alone",
                        "traffic": None},
           "config": {"workload": "100k x 100k 128-D fp32 descriptors, k=2 ratio test "
                                  "(BASELINE.json configs[3]); database rows sharded over the ranks, "
                                  "NCCL all-gather of per-shard top-2 records + merge",
                      "sharding": f"{qg} query groups x {n_db_shards} database shards",
                      "matched": int((mh >= 0).sum()), "index_hash": index_hash}}
    if mg is not None:
        mg.set_trace(True)
        match_step()
        torch.cuda.synchronize()
        ph = mg.match_phase_ms()
        t = torch.tensor([ph["shard_scan"], ph["all_gather"], ph["merge"], ph["total"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["phases_ms_max_over_ranks"] = dict(zip(["shard_scan", "all_gather_incl_wait_for_slowest_rank", "merge", "total"], t.tolist()))
        out["entry"] = "nm_mgpu_match_f32 (C-ABI, one process per GPU)"
        mg.close()
    del A, Bs
    # the sizes the reference can address (its arm prints the same list): one GPU
    if world == 1:
        sizes = []
        for nn in MATCH_SIZES:
            Bq = torch.from_numpy(synth.descriptors(nn, 2)).cuda()
            Aq = torch.from_numpy(synth.descriptors(nn, 1, planted_from=Bq.cpu().numpy())).cuda()
            ms = _time_events(torch, lambda: nm.match(Aq, Bq, 0.8), 5, 2)
            sizes.append({"n": nn, "ms": ms, "gpairs_per_s": nn * nn / ms / 1e6})
            del Aq, Bq
        out["sizes"] = sizes
    torch.cuda.empty_cache()
    return out


def stream_leg(nm, torch, dist, sb, stream_np, n_stream, slo, shi, rank, world, barrier, max_over_ranks):
    """BASELINE.json configs[4]: SIFT + consecutive-frame matching + registration on a stream split into contiguous
    frame ranges with one overlap frame (no descriptors cross GPUs).  See niftymatch_b200/dist.py: StreamRegistrar."""
    from niftymatch_b200.dist import StreamRegistrar
    import numpy as np
    base = torch.from_numpy(stream_np).cuda()                   # 32 rendered frames: 2 scenes x 16 drift steps
    reg = StreamRegistrar(sb, chunk=sb.max_batch)

    def frames_of(t0, t1):
        # frame t of the stream = rendered frame t % 32 (scene changes every 16 frames: those pairs do not register)
        idx = torch.arange(t0, t1, device="cuda") % base.shape[0]
        return base.index_select(0, idx)

    n_local = shi - slo
    reg.run(frames_of, slo, min(shi, slo + sb.max_batch + 1))   # warm-up: one chunk
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    Hs, st = reg.run(frames_of, slo, shi)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    Hh, sth = Hs.cpu().numpy(), st.cpu().numpy()
    # order-independent digest of (pair index, homography bits, inliers): equal for every sharding of the same stream
    own_hi = shi - 1 if rank + 1 < world else shi - 1            # pairs [slo, shi - 1)
    pair_idx = np.arange(slo, slo + Hh.shape[0], dtype=np.uint64)
    words = np.concatenate([Hh.view(np.uint32).astype(np.uint64), sth[:, 1:2].astype(np.uint64)], axis=1)
    dig = np.bitwise_xor.reduce(((words * np.uint64(0x9E3779B97F4A7C15)) ^ (pair_idx[:, None] * np.uint64(0xBF58476D1CE4E5B9))
                                 ).reshape(-1)) if Hh.size else np.uint64(0)
    t = torch.tensor([int(dig) & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device="cuda")
    ok = torch.tensor([int((sth[:, 0] == 1).sum())], dtype=torch.int64, device="cuda")
    if world > 1:
        gl = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gl, t)
        digest = 0
        for g in gl:
            digest ^= int(g.item())
        dist.all_reduce(ok)
    else:
        digest = int(t.item())
    return {"metric": "stream_frames_per_s_1080p", "value": n_stream / ms * 1e3, "unit": "frames/s", "ms": ms, "scaling": "strong",
            "frames": n_stream, "pairs_registered": int(ok.item()), "homography_digest": f"{digest:016x}",
            "config": {"workload": f"{n_stream}-frame synthetic 1080p stream (BASELINE.json configs[4] asks for 10k: the stream length "
                                   "is NM_BENCH_STREAM_FRAMES, default 2048, to keep the default run within minutes), SIFT + "
                                   "match(t -> t+1) + align_points + batched homography RANSAC; contiguous ranges per GPU, one overlap frame, "
                                   "no collective", "frames_this_rank": n_local}}


if __name__ == "__main__":
    main()
