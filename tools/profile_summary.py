"""Turns one GPU session's ncu output (gpurun_out/) into the tracked summaries under profiles/:

    python tools/profile_summary.py r01            # reads gpurun_out/{launches.csv,prof_*.ncu-rep,bench.json}

writes profiles/<tag>_launches.csv      the ncu launch list (gpu__time_duration per launch, --clock-control none)
       profiles/<tag>_launch_summary.txt per-kernel count / average / share of the step
       profiles/<tag>_ncu_<kernel>.txt  key metrics + top stall reasons of the `ncu --set full` capture
       profiles/<tag>_bench.json        the bench line of the same session (NOT taken under the profiler)
       profiles/ncu_summary.json        per-kernel DRAM bytes per launch (bench.py reads `traffic` from here)
Runs in the dev container (no GPU needed: `ncu -i` only reads the report)."""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
src = os.path.join(ROOT, "gpurun_out")
dst = os.path.join(ROOT, "profiles")
os.makedirs(dst, exist_ok=True)


def short(name):
    return name.split("(")[0].replace("void ", "").strip()


# ---- launch list
lines = [ln for ln in open(os.path.join(src, "launches.csv")) if ln.startswith('"')]
with open(os.path.join(dst, f"{tag}_launches.csv"), "w") as f:
    f.writelines(lines)
agg = collections.OrderedDict()
for r in csv.DictReader(io.StringIO("".join(lines))):
    if r["Metric Name"] == "gpu__time_duration.sum":
        agg.setdefault(short(r["Kernel Name"]), []).append(float(r["Metric Value"].replace(",", "")) / 1e3)
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(dst, f"{tag}_launch_summary.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none (serialised, cold-cache launches; shares are what matter)\n")
    for k, v in agg.items():
        f.write(f"{k:28s} launches={len(v):4d}  avg={sum(v) / len(v):8.1f} us  min={min(v):8.1f}  max={max(v):8.1f}  share={sum(v) / tot:6.1%}\n")
print(open(os.path.join(dst, f"{tag}_launch_summary.txt")).read())

# ---- full captures
summary_path = os.path.join(dst, "ncu_summary.json")
summary = json.load(open(summary_path)) if os.path.exists(summary_path) else {}
for rep in sorted(os.listdir(src)):
    if not rep.endswith(".ncu-rep"):
        continue
    kern = rep[len("prof_"):-len(".ncu-rep")]
    text = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_extract.py"), os.path.join(src, rep)],
                          capture_output=True, text=True).stdout
    with open(os.path.join(dst, f"{tag}_ncu_{kern}.txt"), "w") as f:
        f.write(f"ncu --set full --clock-control none --import-source on, kernel regex k_{kern}; extracted by tools/ncu_extract.py\n" + text)
    raw = subprocess.run(["ncu", "-i", os.path.join(src, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]

    def val(row, name):
        i = hdr.index(name)
        v = float(row[i].replace(",", ""))
        u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1)
    per = [(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"), val(r, "gpu__time_duration.sum")) for r in rows[2:]]
    entry = {"dram_bytes_per_launch": sum(p[0] for p in per) / len(per), "duration_us": sum(p[1] for p in per) / len(per),
             "launches_captured": len(per), "session": tag}
    # rays traced by exactly those launches (ADAPT_ITER_LOG of the same command; the capture skips the first NCU_SKIP launches of the kernel)
    log = os.path.join(src, f"iter_log_{kern}.txt")
    if os.path.exists(log):
        skip = int(os.environ.get("NCU_SKIP", "6"))
        it = [tuple(int(x) for x in ln.split()) for ln in open(log) if ln.strip() and not ln.startswith("#")]
        sel = [r for r in it if skip <= r[0] < skip + len(per)]
        if sel:
            entry["rays_in_launch"] = sum(r[1] for r in sel) / len(sel)
            entry["shadow_rays_in_launch"] = sum(r[2] for r in sel) / len(sel)
            entry["dram_bytes_per_ray"] = entry["dram_bytes_per_launch"] / max(entry["rays_in_launch"] + entry["shadow_rays_in_launch"], 1.0)
    summary[f"k_{kern}"] = entry
json.dump(summary, open(summary_path, "w"), indent=1)
print(json.dumps(summary, indent=1))
if os.path.exists(os.path.join(src, "bench.json")):
    shutil.copy(os.path.join(src, "bench.json"), os.path.join(dst, f"{tag}_bench.json"))
