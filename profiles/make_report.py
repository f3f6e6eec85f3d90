"""Turns the files a `bash gpurun_helper.sh` run left under gpurun_out/final/ into the committed evidence of a round:
  python profiles/make_report.py r02
    profiles/<r>_ncu_<kernel>.txt      ncu --set full summaries (summarize_ncu.py) of the dominant kernels, bench shapes
    profiles/traffic.json              dram__bytes_read + dram__bytes_write per launch, read by bench.py (roofline.traffic)
    profiles/<r>_launches_default.csv  ncu launch list of the default bench command
    profiles/<r>_final_bench_*.json    bench lines (default = USCKF fleet + nested configs, reference arm)
    profiles/<r>_gpu_tests.log, <r>_smoke.log
    profiles/<r>_table.md              the numbers table quoted in DESIGN.md 6.1"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F = os.path.join(ROOT, "gpurun_out", "final")
P = os.path.join(ROOT, "profiles")
r = sys.argv[1] if len(sys.argv) > 1 else "r02"

traffic = {}
for wl, rep in (("usckf", "ncu_usckf"), ("ukfom", "ncu_ukfom"), ("msckf", "ncu_msckf"), ("fusion", "ncu_fusion")):
    path = os.path.join(F, rep + ".ncu-rep")
    if not os.path.exists(path):
        continue
    txt = subprocess.run([sys.executable, os.path.join(P, "summarize_ncu.py"), path], capture_output=True, text=True).stdout
    out = os.path.join(P, "%s_%s.txt" % (r, rep))
    open(out, "w").write("# ncu --set full --clock-control none, one launch of the dominant kernel inside `python bench.py --workload %s` "
                         "(the bench's own batch size); see gpurun_helper.sh\n" % wl + txt)
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, val = rows[0], rows[1], rows[2]

    def get(name):
        i = hdr.index(name)
        v = float(val[i].replace(",", ""))
        return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    traffic[wl] = {"bytes_per_launch": get("dram__bytes_read.sum") + get("dram__bytes_write.sum"),
                   "read": get("dram__bytes_read.sum"), "write": get("dram__bytes_write.sum"),
                   "kernel": val[hdr.index("Kernel Name")], "grid": val[hdr.index("launch__grid_size")],
                   "source": "profiles/%s_%s.txt (ncu --set full of the bench command, bench batch size)" % (r, rep)}
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)

for src, dst in (("launches_default.csv", "%s_launches_default.csv"), ("gpu_tests.log", "%s_gpu_tests.log"), ("smoke.log", "%s_smoke.log"),
                 ("bench_default.json", "%s_final_bench_default.json"), ("bench_reference.json", "%s_final_bench_reference.json"),
                 ("bench_ekf.json", "%s_final_bench_ekf.json"), ("bench_msckf_ekf.json", "%s_final_bench_msckf_ekf.json"),
                 ("bench_safefusion.json", "%s_final_bench_safefusion.json"), ("bench_deadreckon.json", "%s_final_bench_deadreckon.json")):
    if os.path.exists(os.path.join(F, src)):
        shutil.copy(os.path.join(F, src), os.path.join(P, dst % r))

# ---- the table of DESIGN.md 6.1 ------------------------------------------------------------------------------------
d = json.load(open(os.path.join(F, "bench_default.json")))
rows = [("usckf (default) configs[3]/[0], 524 288 inst./GPU", d)] + [("%s, nested" % k, v) for k, v in d["also"].items()]
for w in ("ekf", "msckf_ekf", "safefusion", "deadreckon"):
    p = os.path.join(F, "bench_%s.json" % w)
    if os.path.exists(p):
        rows.append(("%s (SURVEY 8f)" % w, json.load(open(p))))
lines = ["| workload | device-resident | ms/step | binding roofline (frac) | other roofline (frac) | dram traffic / algorithmic | e2e (host buffers) | CPU oracle, %d threads |" % d["cpu_baseline"]["cores"],
         "|---|---|---|---|---|---|---|---|"]
for name, v in rows:
    ro, oth = v["roofline"], v["roofline_other"]
    alg = (ro if ro["bound"] == "hbm" else oth)["algorithmic_bytes_per_unit"] * v["value"] * v["ms_per_step"] * 1e-3
    key = name.split()[0].rstrip(",")
    tr = "%.2f" % (traffic[key]["bytes_per_launch"] / alg) if key in traffic else "n/a"
    cb = v.get("cpu_baseline", {}).get("value")
    lines.append("| %s | %.3g %s | %.3f | %s %.3f | %s %.3f | %s | %.3g | %s |" % (
        name, v["value"], v["unit"].replace("filter-", ""), v["ms_per_step"], ro["bound"], ro["frac"], oth["bound"], oth["frac"], tr,
        v["e2e"]["value"], "%.3g" % cb if cb else "n/a"))
open(os.path.join(P, "%s_table.md" % r), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
