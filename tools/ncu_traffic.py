#!/usr/bin/env python
"""profiles/traffic_rNN.json from a full ncu capture of bench.py's default workload: per-SAMPLE DRAM bytes, L2 bytes and
warp instructions of the two march kernels (bench.py scales them by the run's sample count for its roofline block).

usage: ncu_traffic.py capture.ncu-rep bench_line.json out.json "source description"
The bench line must come from the SAME command that was captured (it carries the samples per launch, the layout and the TF)."""
import csv, hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, line_path, out_path, source = sys.argv[1:5]
line = json.loads([l for l in open(line_path) if l.startswith("{")][-1])
samples = line["config"]["active_samples_per_step_per_gpu"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
units = rows[1]


def val(r, name):
    v = float(r[ix[name]].replace(",", ""))
    u = units[ix[name]]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)


h = hashlib.sha256()
d = os.path.join(ROOT, "differender_b200", "csrc")
for f in sorted(os.listdir(d)):
    if f in ("dr_math.cuh", "dr_kernels.cuh", "dr_desc.h"):      # the sources of the two march kernels
        h.update(open(os.path.join(d, f), "rb").read())
kern = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    key = "fwd_kernel" if "fwd_kernel" in name else ("bwd_kernel" if "bwd_kernel" in name else None)
    if key is None or key in kern:
        continue
    wi, ratio = val(r, "smsp__inst_executed.sum"), val(r, "smsp__thread_inst_executed_per_inst_executed.ratio")
    kern[key] = {
        "kernel": name[:80], "ncu_duration_ms": val(r, "gpu__time_duration.sum") * (1e-3 if units[ix["gpu__time_duration.sum"]] == "us" else 1.0),
        "dram_bytes_per_sample": (val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")) / samples,
        "lts_bytes_per_sample": 32.0 * val(r, "lts__t_sectors.sum") / samples,
        "warp_inst_per_sample": wi / samples, "thread_inst_per_sample": wi * ratio / samples, "avg_active_lanes": ratio,
        "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "l1tex_throughput_pct": val(r, "l1tex__throughput.avg.pct_of_peak_sustained_active"),
        "lts_throughput_pct": val(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        "dram_throughput_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "registers": val(r, "launch__registers_per_thread"),
    }
out = {"source": source, "config": "c3" if "C3" in line["config"]["workload"] else line["config"]["workload"][:2].lower(),
       "views_per_gpu": line["config"]["views_per_gpu"], "layout": line["config"]["volume_layout"], "tf": line["config"]["tf"],
       "samples_per_launch": samples, "csrc_digest": h.hexdigest()[:16], "kernels": kern}
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out, indent=1))
