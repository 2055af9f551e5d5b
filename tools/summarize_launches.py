"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum,...] --csv` launch list
per kernel: count, total time, share and -- when the DRAM counters were collected -- DRAM traffic and bandwidth.

usage: python tools/summarize_launches.py gpurun_out/launches.csv [--json out.json] [> profiles/rNN_launches.md]
"""
import collections
import csv
import json
import re
import sys


def family(name):
	if "umma_gemm" in name or "umma_halo" in name:
		return "gemm"
	if "bn_bwd" in name:
		return "bn_bwd"
	if name.startswith("bn_"):
		return "bn_fwd"
	if "pool" in name:
		return "pool"
	if "ew_kernel" in name or "fill" in name or "cast" in name:
		return "eltwise"
	return "other"


def main(path, jsonpath=None):
	with open(path) as f:
		lines = [line for line in f if line.startswith('"')]
	launches = collections.OrderedDict()
	for row in csv.DictReader(lines):
		name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "").replace("pzumma::", "")
		rec = launches.setdefault(row["ID"], {"name": name})
		v = float(row["Metric Value"].replace(",", ""))
		unit = row["Metric Unit"]
		metric = row["Metric Name"]
		if metric == "gpu__time_duration.sum":
			rec["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
		elif metric.startswith("dram__bytes"):
			rec["dram"] = rec.get("dram", 0.0) + v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
		elif metric.startswith("sm__pipe_tensor"):
			rec["tensor"] = v

	agg = collections.defaultdict(lambda: {"n": 0, "us": 0.0, "dram": 0.0, "tensor_us": 0.0})
	fam = collections.defaultdict(lambda: {"n": 0, "us": 0.0, "dram": 0.0})
	for rec in launches.values():
		if "us" not in rec:
			continue
		famname = family(rec["name"])
		# the engine's launches split like bench.py's families: a launch that moves more than 1 DRAM byte per 108 tensor flops
		# cannot be told from the name, so the split uses what ncu measured -- DRAM GB/s above 1 TB/s with the tensor pipe under
		# 30 % active is an HBM-bound launch (1x1 convolutions), the rest is tensor-pipe work
		if famname == "gemm" and rec.get("tensor", 100.0) < 30.0 and rec.get("dram", 0.0) / max(rec["us"], 1e-9) / 1e3 > 1000.0:
			famname = "gemm_hbm"
		for table, key in ((agg, rec["name"]), (fam, famname)):
			a = table[key]
			a["n"] += 1
			a["us"] += rec["us"]
			a["dram"] += rec.get("dram", 0.0)
			if "tensor_us" in a:
				a["tensor_us"] += rec.get("tensor", 0.0) * rec["us"]
	total = sum(a["us"] for a in agg.values())
	print("# %s: %d launches, %.2f ms of kernel time (serialised, cold-cache ncu replay -- shares, not absolutes)\n" %
		  (path, sum(a["n"] for a in agg.values()), total / 1e3))
	print("| kernel | launches | total us | avg us | share | DRAM MB / launch | DRAM GB/s | tensor pipe active |")
	print("|---|---:|---:|---:|---:|---:|---:|---:|")
	for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
		print("| `%s` | %d | %.1f | %.1f | %.1f%% | %.1f | %.0f | %.1f%% |" % (
			name, a["n"], a["us"], a["us"] / a["n"], 100.0 * a["us"] / total, a["dram"] / a["n"] / 1e6, a["dram"] / a["us"] / 1e3,
			a["tensor_us"] / a["us"]))
	print("\n| family | launches | total us | share | DRAM MB / launch | DRAM GB/s |")
	print("|---|---:|---:|---:|---:|---:|")
	for name, a in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
		print("| %s | %d | %.1f | %.1f%% | %.1f | %.0f |" % (name, a["n"], a["us"], 100.0 * a["us"] / total, a["dram"] / a["n"] / 1e6,
														   a["dram"] / a["us"] / 1e3))
	if jsonpath:
		with open(jsonpath, "w") as f:
			json.dump({"source": path, "families": {name: {"launches": a["n"], "us": a["us"], "share": a["us"] / total,
														  "dram_bytes_per_launch": a["dram"] / a["n"]} for name, a in fam.items()}}, f, indent=1)


if __name__ == "__main__":
	main(sys.argv[1], sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None)
