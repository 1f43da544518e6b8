"""Average DRAM traffic / duration per launch from an ncu report -> profiles/gemm_traffic.json (read by bench.py)."""
import csv, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h, units, data = rows[0], rows[1], rows[2:]
def col(name):
    i = h.index(name)
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(units[i], 1)
    return [float(r[i].replace(",", "")) * scale for r in data]
rd, wr, du = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
names = [r[h.index("Kernel Name")] for r in data]
tens = col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in h else None
res = {"launches": len(data), "dram_bytes_per_launch": (sum(rd) + sum(wr)) / len(data),
       "dram_read_bytes_per_launch": sum(rd) / len(data), "dram_write_bytes_per_launch": sum(wr) / len(data),
       "us_per_launch_under_ncu": sum(du) / len(data),
       "per_launch": [{"kernel": n[:80], "dram_MB": round((a + b) / 1e6, 1), "us": round(d, 1)} for n, a, b, d in zip(names, rd, wr, du)],
       "source": "ncu --set full --clock-control none, " + rep}
if tens: res["tensor_pipe_active_pct_avg"] = sum(tens) / len(tens)
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k != "per_launch"}))
