"""The drop-in seam: run the UNMODIFIED reference tree (Modules / Containers / Optimizers / Cost / Handlers / Models and its
own Backend/*.py function table) on top of this package.

The reference reaches the GPU through one module only, `PuzzleLib.Cuda.Backend` (`getBackend`, `getDeviceCount`,
reference: Cuda/Backend.py:353-370).  `install()` registers a module of that name whose two functions are
`puzzlelib_b200.backend.getBackend / getDeviceCount` BEFORE anything of `PuzzleLib.Backend` is imported (the function
table is filled at import time, Backend/gpuarray.py:201, Dnn.py:490, Blas.py:128), which is exactly what replacing the
file `PuzzleLib/Cuda/Backend.py` by the 6-line body of INTEGRATION.md section 1 does.  `PuzzleLib.Grid` is aliased to
`puzzlelib_b200.grid` (NCCL instead of the CUDA-IPC star, same `runGrid` / `NodeInfo` surface).

Nothing of the reference is copied into this package: the tree is found on disk (`PUZZLELIB_ROOT`, else
`baseline/_ref` next to this repository, which `baseline/build_ref.py` fills from /root/reference).
"""
import importlib
import os
import sys
import types

_installed = None


def referenceRoot():
	root = os.environ.get("PUZZLELIB_ROOT")
	if root is None:
		root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
	if not os.path.isdir(os.path.join(root, "PuzzleLib")):
		raise ImportError(
			"the reference tree is not at %s/PuzzleLib: run `python baseline/build_ref.py` where /root/reference exists, "
			"or point PUZZLELIB_ROOT at a directory that holds the PuzzleLib package" % root
		)
	return root


def _importable(name):
	try:
		return importlib.util.find_spec(name) is not None
	except (ImportError, ValueError):
		return False


def install(root=None, deviceIdx=None):
	"""Make `import PuzzleLib...` resolve to the reference tree under `root` with this package as its CUDA backend.
	Returns the `PuzzleLib` package.  Must run before the first import of PuzzleLib.Backend / Modules."""
	global _installed
	if _installed is not None:
		return _installed

	if "PuzzleLib.Backend.gpuarray" in sys.modules:
		raise ImportError("PuzzleLib.Backend is already imported: its function table is bound to another backend")

	root = referenceRoot() if root is None else root
	if root not in sys.path:
		sys.path.insert(0, root)

	stubs = os.path.join(root, "stubs")
	if os.path.isdir(stubs) and not all(_importable(name) for name in ("h5py", "colorama", "graphviz")):
		sys.path.append(stubs)      # behind site-packages: a real h5py wins when it exists

	from . import backend as b200, grid

	seam = types.ModuleType("PuzzleLib.Cuda.Backend")
	seam.__doc__ = "puzzlelib_b200 behind the reference's Cuda/Backend.py seam"
	seam.getBackend = b200.getBackend
	seam.getDeviceCount = b200.getDeviceCount
	seam.B200Backend = b200.B200Backend

	import PuzzleLib
	import PuzzleLib.Cuda as cudapkg

	sys.modules["PuzzleLib.Cuda.Backend"] = seam
	cudapkg.Backend = seam

	sys.modules["PuzzleLib.Grid"] = grid
	PuzzleLib.Grid = grid

	from PuzzleLib import Config
	Config.backend = Config.Backend.cuda
	if deviceIdx is not None:
		Config.deviceIdx = deviceIdx

	_installed = PuzzleLib
	return PuzzleLib


def calcMode(net, dtype):
	"""`net.calcMode(dtype)` for a calculation type the reference's modules do not whitelist (bfloat16).

	Modules that own parameters convert them through their own calcMode (ConvND.py:106-122, Linear.py:89-103: any dtype the
	gpuarray can `astype` to); modules whose calcMode only checks `T in {float16, float32}` (BatchNormND.py:102-110, pooling,
	activations, ...) have their `calctype` set directly -- their kernels take the type from the arrays they are handed."""
	from PuzzleLib.Containers.Container import Container
	from PuzzleLib.Modules.Module import ModuleError

	for mod in net.modules.values():
		if isinstance(mod, Container):
			calcMode(mod, dtype)
			continue
		try:
			mod.calcMode(dtype)
		except ModuleError:
			mod.calctype = dtype
	net.calctype = dtype
