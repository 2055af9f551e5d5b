"""Top stall sites of one kernel from `ncu -i X.ncu-rep --page source --csv` (SASS view).

usage: ncu -i X.ncu-rep --page source --csv > /tmp/src.csv ; python tools/ncu_hot.py /tmp/src.csv [N]
"""
import csv
import sys


def main(path, top=30):
	rows = list(csv.reader(open(path)))
	hdr = rows[1]
	col = {name: i for i, name in enumerate(hdr)}
	stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
	data = []
	total = 0
	for idx, r in enumerate(rows[2:]):
		if len(r) < len(hdr):
			continue
		n = int(r[col["# Samples"]] or 0)
		total += n
		data.append((n, idx, r))
	print("total samples", total)
	agg = {s: 0 for s in stalls}
	for n, idx, r in data:
		for s in stalls:
			agg[s] += int(r[col[s]] or 0)
	print("stall mix:", ", ".join("%s %.1f%%" % (s[6:], 100.0 * v / max(1, total)) for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
	for n, idx, r in sorted(data, key=lambda t: -t[0])[:top]:
		why = sorted(((int(r[col[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
		print("%5.1f%%  #%-5d %-70s exec=%-8s %s" % (100.0 * n / max(1, total), idx, r[col["Source"]].strip()[:70], r[col["Instructions Executed"]],
												   " ".join("%s:%d" % (w, c) for c, w in why if c)))


if __name__ == "__main__":
	main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
