"""Scoreboard use of a kernel, from `nvdisasm -c -hex` text: which of the 6 scoreboards every global load signals and
which instructions wait on them.  A wait on a scoreboard waits for EVERY outstanding load that signals it, so register-ring
prefetching only works when the ring slots own distinct scoreboards.

usage: python tools/sass_scoreboards.py kernel.sass [pattern ...]
"""
import re
import sys


def decode(path):
	lines = open(path).read().splitlines()
	out = []
	for i, line in enumerate(lines):
		m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", line)
		if not m or i + 1 >= len(lines):
			continue
		m2 = re.search(r"/\* (0x[0-9a-f]+) \*/", lines[i + 1])
		if not m2:
			continue
		hi = int(m2.group(1), 16)
		ctrl = hi >> 41
		out.append({"addr": m.group(1), "text": " ".join(m.group(2).split()), "stall": ctrl & 0xf, "yield": (ctrl >> 4) & 1,
					"wr": (ctrl >> 5) & 7, "rd": (ctrl >> 8) & 7, "wait": (ctrl >> 11) & 0x3f})
	return out


if __name__ == "__main__":
	pats = sys.argv[2:] or ["LDG", "SYNCS", "LDS", "LDTM"]
	for ins in decode(sys.argv[1]):
		if any(p in ins["text"] for p in pats) or ins["wait"]:
			print("%6s  wr=%s rd=%s wait=%s  %s" % (ins["addr"], ins["wr"] if ins["wr"] != 7 else "-", ins["rd"] if ins["rd"] != 7 else "-",
												  "".join(str(b) for b in range(6) if ins["wait"] >> b & 1) or "-", ins["text"][:90]))
