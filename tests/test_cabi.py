"""CPU tier: the C-ABI shared library loads without a GPU, exports every symbol include/pzb200.h declares, and its
host-only logic (error reporting, pool size classes, argument validation) behaves like the reference's."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pzb200.h")


def declared_functions():
	text = open(HEADER).read()
	text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
	return sorted(set(re.findall(r"\b(pz_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path():
	names = declared_functions()
	for required in ("pz_conv2d_fprop", "pz_conv2d_dgrad", "pz_conv2d_wgrad", "pz_gemm", "pz_bn_fwd_train", "pz_bn_bwd",
					 "pz_pool2d_fwd", "pz_maxpool2d_mask_fwd", "pz_softmax_fwd", "pz_act_fwd", "pz_nccl_allreduce_mean"):
		assert required in names
	assert len(names) >= 70


def test_library_exports_every_declared_symbol():
	from puzzlelib_b200 import driver
	assert driver.MISSING == []
	lib = ctypes.CDLL(driver.LIBPATH)
	missing = [name for name in declared_functions() if not hasattr(lib, name)]
	assert missing == []


def test_ctypes_signatures_cover_the_header():
	from puzzlelib_b200 import driver
	bound = set(driver.EXPORTS)
	assert [name for name in declared_functions() if name not in bound] == []


def test_pool_size_classes_follow_the_reference_bins():
	# two mantissa bits: sizes round up to {4,5,6,7} * 2^e (reference Allocator.c:29-67), slack <= 25 %
	from puzzlelib_b200.driver import MemoryPool
	assert MemoryPool.allocSize(1) == 256 and MemoryPool.allocSize(256) == 256
	for n in [257, 1000, 4096, 4097, 5 << 20, (5 << 20) + 1, 123456789, 3 << 30]:
		sz = MemoryPool.allocSize(n)
		assert sz >= n and sz <= n * 1.25 + 256
		mant = sz >> (sz.bit_length() - 3)
		assert mant in (4, 5, 6, 7) and sz == mant << (sz.bit_length() - 3)
	assert MemoryPool.allocSize(4096) == 4096 and MemoryPool.allocSize(4097) == 5120


def test_value_errors_surface_as_python_exceptions_with_the_c_message():
	from puzzlelib_b200 import driver
	with pytest.raises(ValueError, match="unsupported dtype"):
		driver.check(driver.lib.pz_gemm(driver.PZ_I32, None, None, None, 4, 4, 4, 4, 4, 4, 0, 0, 1.0, 0.0, None, None))
	desc = driver.Conv2dDesc(1, 4, 8, 8, 4, 3, 3, 5, 5, 1, 1, 0, 0, 1, 1, 1)       # P, Q should be 6
	with pytest.raises(ValueError, match="inconsistent"):
		driver.check(driver.lib.pz_conv2d_fprop(driver.PZ_F32, ctypes.byref(desc), None, None, None, None, None))
	with pytest.raises(ValueError, match="unknown activation"):
		driver.check(driver.lib.pz_act_fwd(99, driver.PZ_F32, None, None, 10, 0.0, 0.0, None))
	with pytest.raises(ValueError, match="pool2d"):
		driver.check(driver.lib.pz_pool2d_fwd(driver.PZ_F32, 0, None, None, 4, 2, 2, 1, 1, 3, 3, 1, 1, 0, 0, None))


def test_product_path_does_not_import_the_oracle():
	pkg = os.path.join(ROOT, "puzzlelib_b200")
	for dirpath, _, files in os.walk(pkg):
		for name in files:
			if name.endswith((".py", ".cu", ".cuh", ".h")):
				text = open(os.path.join(dirpath, name)).read()
				assert "oracle" not in text, "%s mentions the oracle" % name


def test_backend_refuses_to_run_without_a_gpu():
	# no CPU fallback: on a box without a CUDA device the backend raises instead of degrading
	from puzzlelib_b200 import driver
	count = ctypes.c_int(0)
	if driver.lib.pz_device_count(ctypes.byref(count)) == 0 and count.value > 0:
		pytest.skip("a GPU is present")
	from puzzlelib_b200.backend import B200Backend
	with pytest.raises(Exception):
		B200Backend(0, 2)


def test_gpuarray_view_algebra_without_device_memory():
	# shape / stride bookkeeping is pure host logic; use a fake Buffer so no allocation happens
	from puzzlelib_b200.gpuarray import GPUArray
	from puzzlelib_b200.driver import Buffer
	buf = Buffer(4 * 2 * 3 * 4 * 5, ptr=1 << 20)
	a = GPUArray((2, 3, 4, 5), np.float32, gpudata=buf)
	assert a.strides == (240, 80, 20, 4) and a.contiguous and a.nbytes == 480
	b = a[1, :, 1:3]
	assert b.shape == (3, 2, 5) and b.strides == (80, 20, 4) and not b.contiguous
	assert b.ptr == (1 << 20) + 240 + 20
	assert a.reshape(6, -1).shape == (6, 20) and a.ravel().shape == (120, )
	assert a[0].reshape(12, 5).ptr == a.ptr
	c = a[:, 1]
	assert c.reshape(2, 20).strides == (240, 4)
	with pytest.raises(ValueError):
		a[:, :, 1:3].reshape(48)
	assert a.view(np.int16).shape == (2, 3, 4, 10)
	assert [x[1:] for x in b._chunks()] == [(2, 20, 20)] * 3 or len(b._chunks()) >= 1
