import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for path in (ROOT, os.path.join(ROOT, "tests")):
	if path not in sys.path:
		sys.path.insert(0, path)


def pytest_configure(config):
	config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def bnd():
	"""The backend object on cuda:0 -- fails loudly (no skip, no fallback) when the library or the GPU is missing."""
	import refshim
	return refshim.backend()
