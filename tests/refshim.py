"""Test-side access to the UNMODIFIED reference tree running on this repository's backend.

`install()` puts baseline/_ref (the copy of /root/reference made by baseline/build_ref.py) on sys.path with
`puzzlelib_b200` behind its Cuda/Backend.py seam (puzzlelib_b200/seam.py).  `modules()` gathers the reference's own
Modules / Containers / gpuarray names in one namespace, so tests read `M.Conv2D`, `M.Sequential`, `M.gpuarray`.
"""
import types

_ns = None


def install():
	from puzzlelib_b200 import seam
	return seam.install()


def backend():
	install()
	from PuzzleLib.Backend import gpuarray, Dnn, Blas, Memory                       # noqa: F401 -- fills the function table
	from PuzzleLib.Backend.Kernels import ElementWise, MatVec, Pool, Costs          # noqa: F401
	return gpuarray.backend


def modules():
	global _ns
	if _ns is None:
		backend()
		import PuzzleLib.Modules as Mods
		from PuzzleLib import Config
		from PuzzleLib.Backend import gpuarray
		from PuzzleLib.Containers import Sequential, Parallel, Graph
		from PuzzleLib.Modules.Module import Module, ModuleError
		from PuzzleLib.Variable import Variable

		_ns = types.SimpleNamespace(**{name: getattr(Mods, name) for name in dir(Mods) if not name.startswith("_")})
		_ns.Sequential, _ns.Parallel, _ns.Graph = Sequential, Parallel, Graph
		_ns.Module, _ns.ModuleError, _ns.gpuarray, _ns.Config, _ns.Variable = Module, ModuleError, gpuarray, Config, Variable
	return _ns


def leaves(net):
	"""every non-container module under `net`, in definition order (Containers/Container.py keeps them in `.modules`)"""
	from PuzzleLib.Containers.Container import Container
	for mod in net.modules.values():
		if isinstance(mod, Container):
			yield from leaves(mod)
		else:
			yield mod
