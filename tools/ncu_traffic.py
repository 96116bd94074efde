"""DRAM traffic per launch of one kernel from an `ncu --metrics ... dram__bytes_read.sum,dram__bytes_write.sum --csv` log.

    python tools/ncu_traffic.py profiles/r02_ncu_unet_pass_kernels.csv conv_halo_kernel > profiles/r02_conv_halo_traffic.json

Takes the LAST half of the matching launches (the log holds two passes; the second one is warm)."""
import csv, json, sys
path, pat = sys.argv[1], sys.argv[2]
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
per = {}
order = []
for r in csv.DictReader(lines):
    if pat not in r["Kernel Name"]:
        continue
    i = r["ID"]
    if i not in per:
        per[i] = {}
        order.append(i)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3}.get(unit, 1.0)
    per[i][r["Metric Name"]] = v * scale
ids = order[len(order) // 2:]
rd = [per[i].get("dram__bytes_read.sum", 0.0) for i in ids]
wr = [per[i].get("dram__bytes_write.sum", 0.0) for i in ids]
us = [per[i].get("gpu__time_duration.sum", 0.0) for i in ids]
tp = [per[i].get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) for i in ids]
n = max(1, len(ids))
print(json.dumps({"kernel": pat, "launches": len(ids), "dram_bytes_per_launch": (sum(rd) + sum(wr)) / n,
                  "dram_read_bytes_per_launch": sum(rd) / n, "dram_write_bytes_per_launch": sum(wr) / n,
                  "avg_duration_us_under_ncu": sum(us) / n, "tensor_pipe_active_pct_avg": sum(tp) / n,
                  "source": f"{path} (ncu --metrics, --clock-control none, tools/one_pass.py 512 2: second pass)"}, indent=1))
