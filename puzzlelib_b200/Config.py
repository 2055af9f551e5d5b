"""Global flags of the host-side mirror (reference: Config.py:16-56).  Set them before the first operator call."""
import logging
import sys

deviceIdx = 0
allowMultiContext = False
systemLog = False
logger = None

libname = "pzb200"

globalEvalMode = False
disableDtypeShapeChecks = False
disableModuleCompatChecks = False
verifyData = False
showWarnings = True

# One-pass Add / Replicate-backward (3 tensor passes) instead of the reference's fill(0) + k axpy passes (2k + 1).
# The value produced is identical ((0 + a) + b); set to False to replay the reference's launch sequence exactly.
fuseAdd = True


def getLogger():
	global logger

	if logger is not None:
		return logger

	logger = logging.getLogger(libname)
	logger.setLevel(logging.DEBUG if systemLog else logging.INFO)

	handler = logging.StreamHandler(stream=sys.stdout)
	handler.setFormatter(logging.Formatter("[%(name)s] %(message)s"))

	logger.addHandler(handler)
	return logger
