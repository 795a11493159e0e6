"""DRAM traffic of ONE profiled launch from an `ncu --set full` report -> the JSON bench.py quotes as roofline.traffic.
  python tools/ncu_traffic.py gpurun_out/prof_fused_d12.ncu-rep --dim 12 --T 10000 --n 133200 --out profiles/r2_traffic_fused_d12.json
The seeds-per-launch of the capture is cross-checked against the report itself: executed warp instructions divided by the
kernel's grid (runs = grid x runs-per-CTA)."""
import argparse, csv, json, subprocess

ap = argparse.ArgumentParser()
ap.add_argument("report")
ap.add_argument("--dim", type=int, required=True)
ap.add_argument("--T", type=int, required=True)
ap.add_argument("--n", type=int, required=True)
ap.add_argument("--runs-per-cta", type=int, default=4)
ap.add_argument("--out", required=True)
a = ap.parse_args()
txt = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2]
get = lambda k: data[hdr.index(k)]
unit = lambda k: units[hdr.index(k)]
def to_bytes(k):
    v, u = float(get(k).replace(",", "")), unit(k).lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
grid = int(get("launch__grid_size").replace(",", ""))
runs = grid * a.runs_per_cta
assert abs(runs - a.n) < a.runs_per_cta, f"capture ran {runs} runs, not {a.n}"
out = {"kernel": get("Kernel Name"), "report": a.report.split("/")[-1], "dim": a.dim, "steps": a.T, "seeds_in_launch": a.n,
       "grid": grid, "duration_ms": float(get("gpu__time_duration.sum").replace(",", "")) * {"ms": 1, "us": 1e-3, "s": 1e3}.get(unit("gpu__time_duration.sum"), 1),
       "dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
       "warp_instructions": float(get("smsp__inst_executed.sum").replace(",", ""))}
out["dram_bytes_per_run"] = (out["dram_bytes_read"] + out["dram_bytes_write"]) / a.n
out["warp_instructions_per_run"] = out["warp_instructions"] / a.n
json.dump(out, open(a.out, "w"), indent=1)
print(json.dumps(out, indent=1))
