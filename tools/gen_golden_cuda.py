"""Golden vectors from the reference's OWN CUDA backend (cuDNN 9 / cuBLAS 12 / NVRTC on the GPU box).

    python tools/gen_golden_cuda.py --impl ref  --out gpurun_out/ref_cuda_ops.npz  [--nets gpurun_out/ref_cuda_nets.npz]
    python tools/gen_golden_cuda.py --impl b200 --out gpurun_out/b200_ops.npz      (same table through this repo, for diffing)

Runs the seeded case table of tests/golden_cases.py through `PuzzleLib.Backend.*` -- the reference's function table --
with the chosen backend underneath and stores every input and output.  The `--impl ref` files are committed under
tests/golden/ (ref_cuda_ops.npz, ref_cuda_nets.npz); the GPU parity tests and the CPU oracle tests are held to them.
Needs baseline/_ref (python baseline/build_ref.py) and a GPU; never imported by the product.
"""
import argparse, json, os, sys, time, traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument("--impl", choices=("ref", "b200"), required=True)
	ap.add_argument("--out", required=True)
	ap.add_argument("--nets", default=None)
	ap.add_argument("--only", default=None)
	ap.add_argument("--side", action="store_true", help="the side-module table (golden_cases.SIDE_CASES -> ref_cuda_side.npz) instead of CASES")
	args = ap.parse_args()

	refroot = os.path.join(ROOT, "baseline", "_ref")
	if args.impl == "b200":
		from puzzlelib_b200 import seam
		seam.install(refroot)
	else:
		sys.path.insert(0, refroot)
		sys.path.append(os.path.join(refroot, "stubs"))

	from PuzzleLib import Config
	Config.showWarnings = False

	import golden_cases
	B = golden_cases.bind()
	print("backend:", type(B.gpuarray.backend).__name__, B.gpuarray.getDeviceName(), flush=True)

	table = golden_cases.SIDE_CASES if args.side else golden_cases.CASES
	names = [n for n in table if args.only is None or any(s in n for s in args.only.split(","))]
	out, errors = {}, {}
	for name in names:
		try:
			out.update(golden_cases.run(B, [name]))
		except BaseException as e:
			errors[name] = "%s: %s | %s" % (type(e).__name__, str(e)[:300], traceback.format_exc().splitlines()[-3].strip())
			print("FAILED", name, errors[name], flush=True)

	os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
	np.savez_compressed(args.out, **out)
	print("wrote %s: %d arrays from %d cases, %d failed" % (args.out, len(out), len(names) - len(errors), len(errors)))

	if args.nets:
		nets = {}
		for name in golden_cases.NETS:
			t0 = time.time()
			try:
				nets.update(golden_cases.runNets(B, [name]))
				print("net %s done in %.1f s" % (name, time.time() - t0), flush=True)
			except BaseException as e:
				errors["net:" + name] = "%s: %s" % (type(e).__name__, str(e)[:300])
				traceback.print_exc()
		np.savez_compressed(args.nets, **nets)

	with open(os.path.splitext(args.out)[0] + "_errors.json", "w") as f:
		json.dump(errors, f, indent=1)


if __name__ == "__main__":
	main()
