"""How often does the reference's Modules/BatchNorm3D.py unit test (np.allclose at atol 1e-8 on unseeded data) pass in one try?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from puzzlelib_b200 import seam
seam.install()
from PuzzleLib import Config
Config.showWarnings = False
import importlib
BatchNorm3D, BatchNorm2D, BatchNorm1D, BatchNorm = (importlib.import_module('PuzzleLib.Modules.' + n) for n in ('BatchNorm3D', 'BatchNorm2D', 'BatchNorm1D', 'BatchNorm'))
for mod in (BatchNorm3D, BatchNorm2D, BatchNorm1D, BatchNorm):
	ok = 0
	where = {}
	for i in range(200):
		try:
			mod.unittest()
			ok += 1
		except AssertionError as e:
			import traceback
			tb = traceback.extract_tb(sys.exc_info()[2])[-1]
			where[tb.lineno] = where.get(tb.lineno, 0) + 1
	print(mod.__name__, "passes", ok, "of 200; failing lines:", where, flush=True)
