"""profiles/ncu_traffic.json from an `ncu --set full` capture of the dominant kernels at the benched chunk size:

    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep 37888 > profiles/ncu_traffic.json

Per kernel family: DRAM bytes (read + write) per launch averaged over the captured launches (one coarse-pass and one
fine-pass launch each), and the ncu unit utilisations bench.py quotes next to its live-timed roofline entries."""
import csv
import io
import json
import subprocess
import sys

rep, rays = sys.argv[1], int(sys.argv[2])
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
units = rows[1]


def num(r, k):
    v = float(r[ix[k]])
    u = units[ix[k]]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)


fam = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("ufo::", "")
    key = "k_gather_tc" if "k_gather_tc" in name else "k_view_tc" if "k_view_tc" in name else "k_ray_tc" if "k_ray_tc" in name else name
    e = fam.setdefault(key, {"launches": 0, "bytes": 0.0, "us": 0.0, "l1": [], "l2": [], "issue": [], "tensor": [], "dram": [], "names": []})
    e["launches"] += 1
    e["bytes"] += num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")
    e["names"].append(name)
    for k, m in (("l1", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"), ("l2", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                 ("issue", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                 ("tensor", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                 ("dram", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")):
        e[k].append(float(r[ix[m]]))
out = {"_note": f"ncu --set full --clock-control none, {rays}-ray chunk (the default), NV=3, 1600x1216: per kernel family the DRAM bytes "
                f"(dram__bytes_read.sum + dram__bytes_write.sum) per launch averaged over the captured coarse-pass and fine-pass launches, "
                f"and the unit utilisations of those launches; source: {rep.split('/')[-1]}",
       "rays_per_launch": rays}
for k, e in fam.items():
    n = e["launches"]
    out[k] = {"launches_captured": n, "kernels": sorted(set(e["names"])), "bytes_per_launch": e["bytes"] / n,
              "l1_throughput_pct": sum(e["l1"]) / n, "l2_throughput_pct": sum(e["l2"]) / n, "issue_active_pct": sum(e["issue"]) / n,
              "tensor_pipe_active_pct": sum(e["tensor"]) / n, "dram_throughput_pct": sum(e["dram"]) / n}
print(json.dumps(out, indent=1))
