"""Summarise the round-2 `ncu --set full` raw pages (gpurun_out/r02_<kernel>_raw.csv, one launch each, captured by
tools/capture_profiles.sh on a 16-frame 1080p run: the octave-0 launch of the second iteration) into a table and
per-stage DRAM traffic.  Usage: python tools/summarize_profiles.py > profiles/r02_kernels.md"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
KEYS = [("gpu__time_duration.sum", "time"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp instr"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]
NAMES = ["blur_stream5", "blur_stream7", "blur_stream8", "blur_stream10", "blur_stream13", "extrema", "refine_list", "kprefine",
         "gradmap", "orient", "describe", "tc_scan"]


def load(name):
    rows = list(csv.reader(open(os.path.join(G, f"r02_{name}_raw.csv"))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for k, _ in KEYS + [("Kernel Name", "")]:
        if k in hdr:
            i = hdr.index(k)
            d[k] = (vals[i], units[i])
    stalls = []
    for i, k in enumerate(hdr):
        if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
            try:
                stalls.append((float(vals[i].replace(",", "")), k.split("stalled_")[1]))
            except ValueError:
                pass
    tot = sum(v for v, _ in stalls) or 1.0
    d["stalls"] = ", ".join(f"{k} {100 * v / tot:.0f}" for v, k in sorted(stalls, reverse=True)[:4])
    return d


def num(x):
    return float(x[0].replace(",", ""))


def to_bytes(x):
    return num(x) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(x[1], 1)


def main():
    print("# Round 2 — `ncu --set full --clock-control none` of the hot kernels (one launch each)\n")
    print("Captured by `tools/capture_profiles.sh` through `gpurun` on a B200: `tools/quick_bench.py 1920 1080 64` (64 frames, "
          "octave-0 launch of the second iteration) and `tools/match_once.py 100000 100000 1` (main scan launch).  "
          "Cold-cache, serialised: the percentages are what to read, not the durations.  Raw pages: `r02_<kernel>_raw.csv`.\n")
    print("| kernel | " + " | ".join(t for _, t in KEYS) + " | top stall reasons (% of samples) |")
    print("|---|" + "---:|" * len(KEYS) + "---|")
    traffic = {}
    limits = {}
    for n in NAMES:
        try:
            d = load(n)
        except FileNotFoundError:
            continue
        cells = []
        for k, _ in KEYS:
            if k not in d:
                cells.append("")
                continue
            v, u = d[k]
            if k.startswith("dram__bytes"):
                cells.append(f"{to_bytes(d[k]) / 1e6:.1f} MB")
            elif k == "gpu__time_duration.sum":
                cells.append(f"{num(d[k]) * {'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(u, 1):.1f} us")
            elif k == "smsp__inst_executed.sum":
                cells.append(f"{num(d[k]) / 1e6:.1f} M")
            else:
                cells.append(f"{num(d[k]):.1f}" if "." in v else v)
        print(f"| `{d['Kernel Name'][0].split('(')[0].replace('<unnamed>::', '').replace('void ', '')}` ({n}) | " + " | ".join(cells) + f" | {d['stalls']} |")
        traffic[n] = to_bytes(d["dram__bytes_read.sum"]) + to_bytes(d["dram__bytes_write.sum"])
        limits[n] = {"issue_slots_busy_pct": round(num(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]), 1),
                     "fma_pipe_pct": round(num(d["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]), 1),
                     "alu_pipe_pct": round(num(d["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]), 1),
                     "dram_busy_pct": round(num(d["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]), 1),
                     "top_stalls": d["stalls"]}
    # per-stage DRAM bytes of a 64-frame step: the captured launch is octave 0 of 64 frames; a stage's six octave launches
    # carry 4/3 of octave 0's pixels; the pyramid adds the base blur (R = 7 again, octave 0 only)
    stage = {}
    blur = [traffic.get(f"blur_stream{r}") for r in (5, 7, 8, 10, 13)]
    if all(b is not None for b in blur):
        stage["pyramid"] = sum(blur) * 4 / 3 + traffic["blur_stream7"]
    if "extrema" in traffic:
        stage["extrema"] = (traffic["extrema"] + traffic.get("refine_list", 0.0)) * 4 / 3
    if "gradmap" in traffic:
        stage["gradient"] = traffic["gradmap"] * 4 / 3
    if "orient" in traffic:
        stage["orientation"] = traffic["orient"]
    if "describe" in traffic:
        stage["descriptor"] = traffic["describe"]
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        old = json.load(open(path))
    except Exception:
        old = {}
    old.update(stage)
    # what ncu says bounds each stage's kernel (one `--set full` capture each; profiles/r02_kernels.md)
    old["ncu_limits"] = {"pyramid": limits.get("blur_stream13"), "extrema": limits.get("extrema"), "gradient": limits.get("gradmap"),
                         "orientation": limits.get("orient"), "descriptor": limits.get("describe")}
    old["r02_note"] = ("pyramid / extrema / gradient / orientation / descriptor: dram__bytes_read.sum + dram__bytes_write.sum of the stage's "
                       "kernels per 64-frame step, from the 64-frame octave-0 ncu captures (x4/3 octaves for the per-octave "
                       "launches; pyramid = five level blurs x4/3 + the base blur); profiles/r02_kernels.md")
    json.dump(old, open(path, "w"), indent=1)
    print("\nPer-stage DRAM traffic of a 64-frame step (scaled from these captures; written to `roofline_traffic.json`): " +
          ", ".join(f"{k} {v / 1e9:.2f} GB" for k, v in stage.items()) + ".")


if __name__ == "__main__":
    main()
