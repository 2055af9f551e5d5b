"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total time, share).

usage: python tools/summarize_launches.py gpurun_out/launches.csv [> profiles/rNN_launches.md]
"""
import collections
import csv
import re
import sys


def main(path):
	with open(path) as f:
		lines = [line for line in f if line.startswith('"')]
	agg = collections.defaultdict(lambda: [0, 0.0])
	for row in csv.DictReader(lines):
		if row.get("Metric Name") != "gpu__time_duration.sum":
			continue
		name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
		v = float(row["Metric Value"].replace(",", ""))
		v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1.0)
		agg[name][0] += 1
		agg[name][1] += v
	total = sum(v[1] for v in agg.values())
	print("# %s: %d launches, %.2f ms of kernel time (serialised, cold-cache ncu replay -- shares, not absolutes)\n" %
		  (path, sum(v[0] for v in agg.values()), total / 1e3))
	print("| kernel | launches | total us | avg us | share |")
	print("|---|---:|---:|---:|---:|")
	for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
		print("| `%s` | %d | %.1f | %.1f | %.1f%% |" % (name, n, us, us / n, 100.0 * us / total))


if __name__ == "__main__":
	main(sys.argv[1])
