"""Single-try pass rate of every reference unit test in the GPU tier's lists (tests/test_gpu_seam_reference_unittests.py): their inputs are
unseeded and compared with np.allclose, so the reference's own Unittester retries up to 20 times; a test that passes rarely per try
is a flake risk for the tier.  usage (GPU box): python tools/ref_unittest_flake.py [tries]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from puzzlelib_b200 import seam
seam.install()
from PuzzleLib import Config
Config.showWarnings = False
import test_gpu_seam_reference_unittests as T
from PuzzleLib.Cuda import Backend
bnd = Backend.getBackend(0, initmode=2)
refroot = os.path.join(ROOT, "baseline", "_ref", "PuzzleLib")
tries = int(sys.argv[1]) if len(sys.argv) > 1 else 10


def rate(fn, limit=20.0):
	ok = n = 0
	t0 = time.time()
	for _ in range(tries):
		n += 1
		try:
			fn()
			ok += 1
		except AssertionError:
			pass
		if time.time() - t0 > limit:
			break
	return ok, n


def unittest_of(name):
	mod = importlib.import_module("PuzzleLib." + name.replace("/", "."))
	def run():
		cwd = os.getcwd()
		os.chdir(os.path.dirname(os.path.join(refroot, name)))
		try:
			mod.unittest()
		finally:
			os.chdir(cwd)
	return run


rows = []
for name in T.MODULES:
	if name in T.KNOWN_REFERENCE_FAILURES:
		continue
	rows.append((name, ) + rate(unittest_of(name)))
bnd.dnn.enableTensorOps(False)
bnd.blas.enableTensorOps(False)
for name in T.TENSOR_CORE:
	rows.append((name + " [exact fp32]", ) + rate(unittest_of(name)))
for name in ("CuDnn", "CuBlas"):
	mod = importlib.import_module("PuzzleLib.Cuda.Wrappers." + name)
	rows.append(("Cuda/Wrappers/%s [exact fp32]" % name, ) + rate(lambda: mod.backendTest(Backend)))
bnd.dnn.enableTensorOps(True)
bnd.blas.enableTensorOps(True)
for name in ("CuDnnNorm", "CuDnnMemory", "CuDnnSpatialTf"):
	mod = importlib.import_module("PuzzleLib.Cuda.Wrappers." + name)
	rows.append(("Cuda/Wrappers/%s" % name, ) + rate(lambda: mod.backendTest(Backend)))
for name, ok, n in sorted(rows, key=lambda r: r[1] / max(1, r[2])):
	if ok < n:
		print("%-50s %d / %d" % (name, ok, n), flush=True)
print("%d tests, %d always passed" % (len(rows), sum(1 for r in rows if r[1] == r[2])))
