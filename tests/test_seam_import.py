"""CPU tier: the UNMODIFIED reference tree imports over this package's seam.

`PuzzleLib.Modules`, `Containers`, `Cost`, `Optimizers`, `Handlers` and the model zoo bind ~100 attributes of the backend
object at import time (Backend/gpuarray.py:60-113, Dnn.py:159-338, Blas.py:43-102, Backend/Kernels/*.py).  No GPU here, so the
device queries are faked in a child process; nothing is launched -- this checks the binding surface only.  The reference's own
unit tests over the seam run in the GPU tier (test_gpu_seam_reference_unittests.py).
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFROOT = os.path.join(ROOT, "baseline", "_ref")

CHILD = r"""
import sys
sys.path.insert(0, %r)
from puzzlelib_b200 import driver
driver.Device.count = staticmethod(lambda: 1)
driver.Device.set = lambda self: self
driver.Device.name = lambda self: "no device (import check)"
driver.Device.computeCapability = lambda self: (10, 0)
driver.Device.smCount = staticmethod(lambda: 148)

from puzzlelib_b200 import seam
seam.install()
import PuzzleLib.Modules, PuzzleLib.Containers, PuzzleLib.Cost, PuzzleLib.Optimizers, PuzzleLib.Handlers
from PuzzleLib.Models.Nets.ResNet import loadResNet
from PuzzleLib.Models.Nets.VGG import loadVGG
from PuzzleLib.Models.Nets.LeNet import loadLeNet
from PuzzleLib.Backend import gpuarray, Dnn, Blas, Memory
from PuzzleLib.Backend.Kernels import ElementWise, MatVec, Pool, Costs, Pad, PRelu, Upsample, Embedder
import PuzzleLib.Grid as Grid

from puzzlelib_b200.backend import B200Backend
from puzzlelib_b200.gpuarray import GPUArray
assert isinstance(gpuarray.backend, B200Backend) and gpuarray.GPUArray is GPUArray
assert Grid.__name__ == "puzzlelib_b200.grid"

# every name of the function table that the reference fills for the CUDA backend is filled
for mod in (gpuarray, Dnn, Blas, Memory, ElementWise, MatVec, Pool, Costs):
	empty = [name for name, val in vars(mod).items() if val is None and not name.startswith("_")]
	empty = [name for name in empty if name not in ("setupDebugAllocator", )]
	assert not empty, (mod.__name__, empty)
print("seam ok")
"""


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFROOT, "PuzzleLib")), reason="baseline/_ref is not built (python baseline/build_ref.py)")
def test_reference_tree_imports_over_the_seam():
	out = subprocess.run([sys.executable, "-c", CHILD % ROOT], capture_output=True, text=True, timeout=300)
	assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
	assert "seam ok" in out.stdout
