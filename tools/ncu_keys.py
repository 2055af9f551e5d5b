"""Key metrics of each kernel in an .ncu-rep (raw page): python tools/ncu_keys.py file.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
		"gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
		"sm__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
		"launch__registers_per_thread", "launch__grid_size", "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
		"lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "l1tex__t_bytes.sum",
		"smsp__inst_executed.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def main(path):
	out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
	rows = list(csv.reader(out.splitlines()))
	hdr, units = rows[0], rows[1]
	for vals in rows[2:]:
		for key in KEYS:
			for i, name in enumerate(hdr):
				if name == key:
					print("%-68s %s %s" % (key, vals[i], units[i]))
		print()


if __name__ == "__main__":
	main(sys.argv[1])
