"""Measure the dense TF32 tensor-core throughput of this GPU with the vendor library (cuBLAS through torch.matmul, fp32 storage,
allow_tf32) -- the roofline denominator BASELINE.md section 2 asks the builder to measure instead of the 0.5 x bf16 proxy.
A yardstick only: nothing of the product path touches torch.  Writes profiles/r02_tf32_peak.json; bench.py reads it."""
import json
import os
import sys
import time

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def measure(n, dtype, tf32, reps):
	torch.backends.cuda.matmul.allow_tf32 = tf32
	a = torch.randn(n, n, device="cuda", dtype=dtype)
	b = torch.randn(n, n, device="cuda", dtype=dtype)
	for _ in range(3):
		a @ b
	torch.cuda.synchronize()
	best, total = 0.0, 0.0
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	for _ in range(reps):
		e0.record()
		a @ b
		e1.record()
		e1.synchronize()
		tf = 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
		best = max(best, tf)
		total += tf
	# sustained: back-to-back launches for about a second
	torch.cuda.synchronize()
	t0 = time.perf_counter()
	count = 0
	e0.record()
	while time.perf_counter() - t0 < 1.0:
		for _ in range(10):
			a @ b
		count += 10
		torch.cuda.synchronize()
	e1.record()
	e1.synchronize()
	sustained = 2.0 * n ** 3 * count / (e0.elapsed_time(e1) * 1e-3) / 1e12
	return best, sustained


def main():
	n = 8192
	tf32_burst, tf32_sustained = measure(n, torch.float32, True, 10)
	bf16_burst, bf16_sustained = measure(n, torch.bfloat16, True, 10)
	out = {"tf32_tflops_burst": tf32_burst, "tf32_tflops_sustained": tf32_sustained, "bf16_tflops_burst": bf16_burst,
		   "bf16_tflops_sustained": bf16_sustained, "how": "torch.matmul %d^3 (cuBLAS), fp32 storage with allow_tf32 / bf16; best of 10 and ~1 s back to back" % n,
		   "device": torch.cuda.get_device_name(0)}
	print(json.dumps(out))
	path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "tf32_peak.json")
	os.makedirs(os.path.dirname(path), exist_ok=True)
	with open(path, "w") as f:
		json.dump(out, f, indent=1)


if __name__ == "__main__":
	main()
