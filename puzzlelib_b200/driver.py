"""ctypes binding of libpzb200.so -- the `Driver`-shaped layer of the backend.

Mirrors the surface of the reference's CPython extension `PuzzleLib.Cuda.Driver`
(reference: Cuda/Source/Core/{Device,Buffer,Allocator,Stream}.c, Driver.h) on top of the plain C-ABI declared in
include/pzb200.h.  There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
"""
import ctypes
import os
from ctypes import c_int, c_int32, c_int64, c_uint8, c_uint16, c_uint32, c_uint64, c_size_t, c_float, c_double, c_void_p, \
	c_char_p, POINTER, byref

import numpy as np

try:
	from ml_dtypes import bfloat16 as _bf16
	bfloat16 = np.dtype(_bf16)
except ImportError:  # pragma: no cover
	bfloat16 = None


class CudaError(Exception):
	pass


class CuDnnError(CudaError):
	pass


class CuBlasError(CudaError):
	pass


class NcclError(CudaError):
	pass


PZ_OK, PZ_ERR_VALUE, PZ_ERR_CUDA, PZ_ERR_MEMORY, PZ_ERR_NCCL, PZ_ERR_UNSUPPORTED = range(6)

LIBNAME = "libpzb200.so"
# PZB200_LIB: an instrumented build of the same library (tools/gpu_timeline.sh); never a different implementation
LIBPATH = os.environ.get("PZB200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), LIBNAME)


def _load():
	if not os.path.exists(LIBPATH):
		raise ImportError(
			"%s is not built: run `python -c 'import __graft_entry__ as g; g.build()'` or "
			"`make -C puzzlelib_b200/csrc` (there is no CPU fallback)" % LIBPATH
		)
	return ctypes.CDLL(LIBPATH, mode=ctypes.RTLD_GLOBAL)


lib = _load()
lib.pz_last_error.restype = c_char_p
lib.pz_launch_count.restype = c_uint64
lib.pz_pool_alloc_size.restype = c_size_t
lib.pz_pool_alloc_size.argtypes = [c_size_t]
lib.pz_conv2d_dgrad_workspace.restype = c_size_t


class Conv2dDesc(ctypes.Structure):
	_fields_ = [(name, c_int) for name in (
		"N", "C", "H", "W", "K", "R", "S", "P", "Q",
		"stride_h", "stride_w", "pad_h", "pad_w", "dil_h", "dil_w", "groups"
	)]


_P = c_void_p
_SIGNATURES = {
	"pz_device_count": [POINTER(c_int)],
	"pz_device_set": [c_int],
	"pz_device_get": [POINTER(c_int)],
	"pz_device_name": [c_int, c_char_p, c_int],
	"pz_device_sm_count": [POINTER(c_int)],
	"pz_device_cc": [POINTER(c_int), POINTER(c_int)],
	"pz_device_synchronize": [],
	"pz_mem_info": [POINTER(c_size_t), POINTER(c_size_t)],
	"pz_malloc": [POINTER(_P), c_size_t],
	"pz_free": [_P],
	"pz_host_alloc": [POINTER(_P), c_size_t],
	"pz_host_free": [_P],
	"pz_pool_create": [POINTER(_P)],
	"pz_pool_destroy": [_P],
	"pz_pool_alloc": [_P, c_size_t, POINTER(_P), POINTER(c_size_t)],
	"pz_pool_release": [_P, _P, c_size_t],
	"pz_pool_free_held": [_P],
	"pz_pool_stats": [_P, POINTER(c_size_t), POINTER(c_size_t), POINTER(c_size_t), POINTER(c_size_t)],
	"pz_memcpy_h2d": [_P, _P, c_size_t, _P, c_int],
	"pz_memcpy_d2h": [_P, _P, c_size_t, _P, c_int],
	"pz_memcpy_d2d": [_P, _P, c_size_t, _P],
	"pz_memcpy2d": [_P, c_size_t, _P, c_size_t, c_size_t, c_size_t, c_int, _P],
	"pz_memset8": [_P, c_uint8, c_size_t, _P],
	"pz_memset16": [_P, c_uint16, c_size_t, _P],
	"pz_memset32": [_P, c_uint32, c_size_t, _P],
	"pz_fill64": [_P, c_uint64, c_int64, _P],
	"pz_set_default_stream": [_P],
	"pz_graph_begin": [_P],
	"pz_graph_end": [_P, POINTER(_P)],
	"pz_graph_launch": [_P, _P],
	"pz_graph_destroy": [_P],
	"pz_debug_timeline": [_P],
	"pz_divnorm_fwd": [c_int, _P, _P, _P, c_int64, c_int64, c_int64, c_int, c_float, c_float, c_float, _P],
	"pz_divnorm_bwd": [c_int, _P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int64, c_int, c_float, c_float, c_float, _P],
	"pz_spatialtf_fwd": [c_int, _P, _P, _P, _P, c_int64, c_int64, c_int, c_int, c_int, c_int, _P],
	"pz_spatialtf_bwd": [c_int, _P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int, c_int, c_int, c_int, _P],
	"pz_ctc_loss": [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P],
	"pz_prelu_fwd": [_P, _P, _P, c_int64, c_int64, c_int64, c_int, _P],
	"pz_prelu_bwd_data": [_P, _P, _P, _P, c_int64, c_int64, c_int64, c_int, _P],
	"pz_prelu_bwd_params": [_P, _P, _P, c_int64, c_int64, c_int64, c_int, _P],
	"pz_reflectpad_fwd": [c_int, _P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, _P],
	"pz_reflectpad_bwd": [c_int, _P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, _P],
	"pz_embed_fwd": [c_int, _P, _P, _P, c_int64, c_int64, _P],
	"pz_embed_bwd": [c_int, _P, _P, _P, c_float, c_int64, c_int64, _P],
	"pz_upsample_nearest_fwd": [_P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, _P],
	"pz_upsample_nearest_bwd": [_P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, _P],
	"pz_upsample_linear_fwd": [_P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_int, _P],
	"pz_upsample_linear_bwd": [_P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_int, _P],
	"pz_stream_create": [POINTER(_P)],
	"pz_stream_destroy": [_P],
	"pz_stream_synchronize": [_P],
	"pz_event_create": [POINTER(_P)],
	"pz_event_create_sync": [POINTER(_P)],
	"pz_event_destroy": [_P],
	"pz_event_record": [_P, _P],
	"pz_event_synchronize": [_P],
	"pz_event_elapsed_ms": [_P, _P, POINTER(c_float)],
	"pz_stream_wait_event": [_P, _P],
	"pz_profile_enable": [c_int],
	"pz_profile_collect": [c_int, POINTER(c_double), POINTER(c_double), POINTER(c_double), POINTER(c_uint64)],
	"pz_act_fwd": [c_int, c_int, _P, _P, c_int64, c_float, c_float, _P],
	"pz_act_bwd": [c_int, c_int, _P, _P, _P, c_int64, c_float, c_float, _P],
	"pz_axpy": [c_int, _P, _P, c_float, c_int64, _P],
	"pz_axpby": [c_int, _P, _P, c_float, _P, c_float, c_int64, _P],
	"pz_axpy2": [c_int, _P, _P, c_float, _P, c_float, c_int64, _P],
	"pz_axpy2_relu": [c_int, _P, _P, _P, c_float, _P, c_float, c_int64, _P],
	"pz_axpy2_relu_bwd": [c_int, _P, _P, _P, c_float, _P, c_float, _P, c_int64, _P],
	"pz_scale_shift": [c_int, _P, _P, c_float, c_float, c_int64, _P],
	"pz_mul": [c_int, _P, _P, _P, c_int64, _P],
	"pz_add2": [c_int, _P, _P, _P, c_int64, _P],
	"pz_cast": [c_int, _P, c_int, _P, c_int64, _P],
	"pz_set_exact_fp32": [c_int],
	"pz_act_fwd_slice": [c_int, c_int, _P, _P, c_int64, c_float, c_float, c_int64, c_int64, c_int64, _P],
	"pz_act_bwd_slice": [c_int, c_int, _P, _P, _P, c_int64, c_float, c_float, c_int64, c_int64, c_int64, _P],
	"pz_axpby_slice": [c_int, _P, _P, c_float, _P, c_float, c_int64, c_int64, c_int64, c_int64, _P],
	"pz_mul_slice": [c_int, _P, _P, _P, c_int64, c_int64, c_int64, c_int64, _P],
	"pz_eltwise": [c_int, c_int, POINTER(_P), c_int, POINTER(c_float), c_int, POINTER(c_int), c_int64, c_int64, c_int64, c_int64, _P],
	"pz_dropout_slice": [c_int, _P, _P, _P, c_uint32, c_float, c_int64, c_int64, c_int64, c_int64, c_int64, _P],
	"pz_sgd_momentum": [c_int, _P, _P, _P, c_float, c_float, c_int64, _P],
	"pz_sgd_nesterov": [c_int, _P, _P, _P, c_float, c_float, c_int64, _P],
	"pz_adam": [c_int, _P, _P, _P, _P, c_float, c_float, c_float, c_float, c_int64, _P],
	"pz_cross_entropy": [_P, _P, _P, c_int64, c_int64, c_int64, _P, _P, _P],
	"pz_count_mismatch": [_P, _P, c_int64, _P, _P],
	"pz_svm": [c_int, _P, _P, c_int64, c_int64, c_int64, _P, _P, _P],
	"pz_cost_reduce": [c_int, _P, _P, _P, c_float, c_int64, _P, _P],
	"pz_reduce_minmax_i32": [_P, c_int64, c_int, _P, _P],
	"pz_permute": [c_int, _P, _P, c_int, _P, _P, _P],
	"pz_matvec": [c_int, _P, _P, _P, c_int64, c_int64, c_int64, c_int, c_float, c_float, _P],
	"pz_vec_reduce": [c_int, c_int, _P, _P, c_int64, _P, _P],
	"pz_lrn_fwd": [c_int, c_int, _P, _P, c_int64, c_int64, c_int64, c_int64, c_int, c_float, c_float, c_float, _P],
	"pz_lrn_bwd": [c_int, c_int, _P, _P, _P, _P, c_int64, c_int64, c_int64, c_int64, c_int, c_float, c_float, c_float, _P],
	"pz_rng_fill": [c_int, _P, c_int64, c_uint64, c_uint64, c_float, c_float, _P],
	"pz_rng_fill_dev": [c_int, _P, c_int64, c_uint64, _P, c_float, c_float, _P],
	"pz_dropout": [c_int, _P, _P, _P, c_uint32, c_float, c_int64, c_int64, _P],
	"pz_reduce_minmax": [c_int, _P, c_int64, c_int, _P, _P],
	"pz_addvec2mat": [c_int, _P, _P, _P, c_int64, c_int64, c_int64, c_int, c_int64, _P],
	"pz_matsum": [c_int, _P, _P, c_int64, c_int64, c_int64, c_int, c_float, c_float, _P],
	"pz_argminmax": [c_int, _P, _P, c_int64, c_int64, c_int64, c_int, _P],
	"pz_bn_fwd_train": [c_int, _P, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P, c_double, c_double, _P],
	"pz_bn_fwd_train_relu": [c_int, _P, _P, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P, c_double, c_double, _P],
	"pz_bn_fwd_infer": [c_int, _P, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P, c_double, _P],
	"pz_bn_bwd": [c_int, _P, _P, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P],
	"pz_bn_bwd_acc": [c_int, _P, _P, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P, c_float, c_float, _P, c_float, c_float, _P],
	"pz_pool2d_fwd": [c_int, c_int, _P, _P, c_int64] + [c_int] * 10 + [_P],
	"pz_pool2d_bwd": [c_int, c_int, _P, _P, _P, _P, c_int64] + [c_int] * 10 + [_P],
	"pz_maxpool2d_mask_fwd": [_P, _P, _P, c_int64] + [c_int] * 10 + [_P],
	"pz_maxpool2d_mask_bwd": [_P, _P, _P, c_int64] + [c_int] * 10 + [_P],
	"pz_maxunpool2d_fwd": [_P, _P, _P, c_int64, c_int, c_int, _P],
	"pz_maxunpool2d_bwd": [_P, _P, _P, c_int64, c_int, c_int, _P],
	"pz_softmax_fwd": [c_int, c_int, _P, _P, c_int64, c_int64, c_int64, _P],
	"pz_softmax_bwd": [c_int, c_int, _P, _P, _P, c_int64, c_int64, c_int64, _P],
	"pz_lstm_cell_fwd": [_P, _P, _P, _P, _P, _P, c_int64, c_int64, _P],
	"pz_lstm_cell_bwd": [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int, _P],
	"pz_gru_cell_fwd": [_P, _P, _P, _P, _P, _P, c_int64, c_int64, _P],
	"pz_gru_cell_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, _P],
	"pz_rnn_cell_fwd": [_P, _P, _P, c_int64, c_int64, c_int, _P],
	"pz_rnn_cell_bwd": [_P, _P, _P, _P, c_int64, c_int, _P],
	"pz_add2d": [c_int, _P, c_int64, _P, c_int64, c_int64, c_int64, _P],
	"pz_gemm": [c_int, _P, _P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_int, c_float, c_float,
				_P, _P],
	"pz_conv2d_fprop": [c_int, POINTER(Conv2dDesc), _P, _P, _P, _P, _P],
	"pz_conv2d_dgrad_workspace": [c_int, POINTER(Conv2dDesc)],
	"pz_conv2d_dgrad": [c_int, POINTER(Conv2dDesc), _P, _P, _P, _P, _P, c_size_t, _P],
	"pz_conv2d_wgrad": [c_int, POINTER(Conv2dDesc), _P, _P, _P, c_float, c_float, _P],
	"pz_bias_grad": [c_int, _P, _P, c_int64, c_int64, c_int64, c_float, c_float, _P],
	"pz_nccl_version": [POINTER(c_int)],
	"pz_nccl_unique_id": [_P],
	"pz_mean_sgd_momentum": [c_int, _P, _P, _P, c_int64, c_float, c_float, c_float, _P],
	"pz_nccl_comm_init": [POINTER(_P), c_int, c_int, _P],
	"pz_nccl_comm_destroy": [_P],
	"pz_nccl_allreduce_mean": [_P, c_int, _P, c_int64, c_float, _P],
	"pz_nccl_broadcast": [_P, c_int, _P, c_int64, c_int, _P],
	"pz_nccl_allreduce_avg_segments": [_P, c_int, POINTER(_P), POINTER(c_int64), c_int, _P],
	"pz_nccl_allreduce_sgd_momentum": [_P, c_int, _P, _P, _P, c_int64, c_float, c_float, c_float, _P],
}

EXPORTS = sorted(_SIGNATURES) + ["pz_last_error", "pz_version", "pz_launch_count", "pz_pool_alloc_size", "pz_exact_fp32"]


def _bind():
	missing = []
	for name, argtypes in _SIGNATURES.items():
		try:
			fn = getattr(lib, name)
		except AttributeError:
			missing.append(name)
			continue
		fn.argtypes = argtypes
		if name != "pz_conv2d_dgrad_workspace":
			fn.restype = c_int
	return missing


MISSING = _bind()


def raiseOnStatus(status, errtype=CudaError):
	if status == PZ_OK:
		return

	msg = lib.pz_last_error().decode("utf-8", "replace")

	if status == PZ_ERR_VALUE:
		raise ValueError(msg)
	elif status == PZ_ERR_MEMORY:
		raise MemoryError(msg)
	elif status == PZ_ERR_UNSUPPORTED:
		raise NotImplementedError(msg)
	elif status == PZ_ERR_NCCL:
		raise NcclError(msg)

	raise errtype(msg)


def check(status):
	if status:
		raiseOnStatus(status)


# --------------------------------------------------------------------------------------------------------- deferred fill
# `y.fill(0)` followed by `y += a * x` (twice) is how the reference's Add / Replicate modules build a sum (Modules/Add.py:15-23,
# Replicate.py:18-29): 7 passes over tensors of the largest activations.  The backend keeps at most ONE such launch pending --
# the zero fill, then the scaled copy it turns into -- and issues it fused with the next accumulation into the same array (3
# passes, same bits).  Every access to a device pointer (`GPUArray.ptr`), every synchronisation and every raw Buffer
# operation flushes the pending launch first, so nothing else can observe the array before it is materialised.
deferred = None


def flushDeferred():
	global deferred
	if deferred is not None:
		op, deferred = deferred, None
		op.flush()


# --------------------------------------------------------------------------------------------------------- dtypes
PZ_F32, PZ_F16, PZ_BF16, PZ_F64, PZ_I8, PZ_U8, PZ_I16, PZ_U16, PZ_I32, PZ_U32, PZ_I64, PZ_U64 = range(12)

_DTYPE_CODES = {
	np.dtype(np.float32): PZ_F32, np.dtype(np.float16): PZ_F16, np.dtype(np.float64): PZ_F64,
	np.dtype(np.int8): PZ_I8, np.dtype(np.uint8): PZ_U8, np.dtype(np.int16): PZ_I16, np.dtype(np.uint16): PZ_U16,
	np.dtype(np.int32): PZ_I32, np.dtype(np.uint32): PZ_U32, np.dtype(np.int64): PZ_I64, np.dtype(np.uint64): PZ_U64,
}
if bfloat16 is not None:
	_DTYPE_CODES[bfloat16] = PZ_BF16


def dtypeCode(dtype):
	try:
		return _DTYPE_CODES[dtype]
	except (KeyError, TypeError):
		try:
			return _DTYPE_CODES[np.dtype(dtype)]
		except KeyError:
			raise TypeError("unsupported dtype %s" % dtype)


# --------------------------------------------------------------------------------------------------------- device
class Device:
	"""reference: Cuda/Source/Core/Device.c"""

	def __init__(self, index=0):
		self.index = index

	@staticmethod
	def count():
		n = c_int(0)
		check(lib.pz_device_count(byref(n)))
		return n.value

	@staticmethod
	def getCurrent():
		n = c_int(0)
		check(lib.pz_device_get(byref(n)))
		return Device(n.value)

	def set(self):
		flushDeferred()               # a pending launch belongs to the device that was current when it was held back
		check(lib.pz_device_set(self.index))
		return self

	def name(self):
		buf = ctypes.create_string_buffer(256)
		check(lib.pz_device_name(self.index, buf, 256))
		return buf.value.decode()

	def computeCapability(self):
		major, minor = c_int(0), c_int(0)
		check(lib.pz_device_cc(byref(major), byref(minor)))
		return major.value, minor.value

	@staticmethod
	def smCount():
		n = c_int(0)
		check(lib.pz_device_sm_count(byref(n)))
		return n.value

	@staticmethod
	def synchronize():
		flushDeferred()
		check(lib.pz_device_synchronize())


def getMemoryInfo():
	free, total = c_size_t(0), c_size_t(0)
	check(lib.pz_mem_info(byref(free), byref(total)))
	return free.value, total.value


def launchCount():
	return int(lib.pz_launch_count())


PROF_FAMILIES = {"gemm": 0, "bn_fwd": 1, "bn_bwd": 2, "eltwise": 3, "pool": 4, "other": 5, "gemm_hbm": 6}


def profileEnable(on):
	check(lib.pz_profile_enable(1 if on else 0))


def profileCollect(family):
	ms, flops, nbytes, launches = c_double(0), c_double(0), c_double(0), c_uint64(0)
	check(lib.pz_profile_collect(PROF_FAMILIES[family] if isinstance(family, str) else family, byref(ms), byref(flops),
								 byref(nbytes), byref(launches)))
	return {"ms": ms.value, "flops": flops.value, "bytes": nbytes.value, "launches": launches.value}


# --------------------------------------------------------------------------------------------------------- memory
class MemoryPool:
	"""Binned caching allocator living in the C library (reference: Cuda/Source/Core/Allocator.c)."""
	__slots__ = ["handle", "holding", "__weakref__"]

	def __init__(self):
		h = c_void_p()
		check(lib.pz_pool_create(byref(h)))
		self.handle = h.value
		self.holding = True

	def allocate(self, nbytes):
		return Buffer(nbytes, allocator=self)

	def freeHeld(self):
		check(lib.pz_pool_free_held(self.handle))

	def stopHolding(self):
		self.holding = False
		self.freeHeld()

	def getStats(self):
		hb, hbytes, ab, abytes = c_size_t(0), c_size_t(0), c_size_t(0), c_size_t(0)
		check(lib.pz_pool_stats(self.handle, byref(hb), byref(hbytes), byref(ab), byref(abytes)))
		return {"heldBlocks": hb.value, "heldBytes": hbytes.value, "activeBlocks": ab.value, "activeBytes": abytes.value}

	@staticmethod
	def allocSize(nbytes):
		return int(lib.pz_pool_alloc_size(nbytes))

	def __del__(self):
		try:
			if self.handle is not None and lib is not None:
				lib.pz_pool_destroy(self.handle)
				self.handle = None
		except Exception:
			pass


_pool_alloc = lib.pz_pool_alloc
_pool_release = lib.pz_pool_release


class Buffer:
	"""Owning (or borrowed-slice) range of device memory (reference: Cuda/Source/Core/Buffer.c, Driver.h:48-60)."""
	__slots__ = ["ptr", "size", "parent", "pool", "granted", "owner"]

	def __init__(self, nbytes=0, allocator=None, ptr=None, parent=None):
		self.parent = parent
		self.pool = None
		self.granted = 0
		self.owner = False

		if ptr is not None:
			self.ptr, self.size = ptr, nbytes
			return

		self.size = nbytes
		p = c_void_p()

		if allocator is not None:
			granted = c_size_t(0)
			st = _pool_alloc(allocator.handle, nbytes, byref(p), byref(granted))
			if st:
				raiseOnStatus(st)
			self.pool, self.granted = allocator, granted.value
		else:
			st = lib.pz_malloc(byref(p), nbytes)
			if st:
				raiseOnStatus(st)

		self.ptr = p.value or 0
		self.owner = True

	def __getitem__(self, item):
		if not isinstance(item, slice) or item.step not in (None, 1):
			raise ValueError("buffer slices must be contiguous byte ranges")
		start, stop, _ = item.indices(self.size)
		if stop < start:
			raise ValueError("invalid buffer slice")
		return Buffer(stop - start, ptr=self.ptr + start, parent=self)

	@property
	def int_ptr(self):
		return self.ptr

	def fillD8(self, value, stream=None):
		flushDeferred()
		check(lib.pz_memset8(self.ptr, value & 0xff, self.size, stream))

	def fillD16(self, value, stream=None):
		flushDeferred()
		check(lib.pz_memset16(self.ptr, value & 0xffff, self.size // 2, stream))

	def fillD32(self, value, stream=None):
		flushDeferred()
		check(lib.pz_memset32(self.ptr, value & 0xffffffff, self.size // 4, stream))

	def copy(self, dst=None, allocator=None, stream=None):
		flushDeferred()
		if dst is None:
			dst = Buffer(self.size, allocator=allocator)
		elif dst.size < self.size:
			raise ValueError("destination buffer is too small")
		check(lib.pz_memcpy_d2d(dst.ptr, self.ptr, self.size, stream))
		return dst

	def set(self, host, stream=None):
		flushDeferred()
		host = np.ascontiguousarray(host)
		if host.nbytes > self.size:
			raise ValueError("host array is larger than the buffer")
		check(lib.pz_memcpy_h2d(self.ptr, host.ctypes.data, host.nbytes, stream, 0))

	def get(self, host, stream=None):
		flushDeferred()
		check(lib.pz_memcpy_d2h(host.ctypes.data, self.ptr, min(host.nbytes, self.size), stream, 0))
		return host

	def free(self):
		if self.owner and self.ptr:
			if self.pool is not None and self.pool.holding and self.pool.handle is not None:
				_pool_release(self.pool.handle, self.ptr, self.granted)
			else:
				lib.pz_free(self.ptr)
		self.owner = False
		self.ptr = 0
		self.parent = None

	def __del__(self):
		try:
			self.free()
		except Exception:
			pass


class PinnedBuffer:
	"""Page-locked host staging memory exposed as a numpy array (used by the e2e bench path)."""

	def __init__(self, shape, dtype):
		dtype = np.dtype(dtype)
		nbytes = int(np.prod(shape)) * dtype.itemsize
		p = c_void_p()
		check(lib.pz_host_alloc(byref(p), nbytes))
		self.ptr = p.value
		raw = (ctypes.c_uint8 * nbytes).from_address(self.ptr)
		self.array = np.frombuffer(raw, dtype=dtype).reshape(shape)

	def free(self):
		if self.ptr:
			self.array = None
			lib.pz_host_free(self.ptr)
			self.ptr = None

	def __del__(self):
		try:
			self.free()
		except Exception:
			pass


# --------------------------------------------------------------------------------------------------------- streams
class Stream:
	def __init__(self):
		h = c_void_p()
		check(lib.pz_stream_create(byref(h)))
		self.handle = h.value

	def waitEvent(self, event):
		check(lib.pz_stream_wait_event(self.handle, event.handle))

	def synchronize(self):
		flushDeferred()
		check(lib.pz_stream_synchronize(self.handle))

	def __del__(self):
		try:
			if self.handle:
				lib.pz_stream_destroy(self.handle)
				self.handle = None
		except Exception:
			pass


# scalars that change from step to step (batch-norm running-average factor, bias-corrected learning rates ...) are frozen into
# a captured graph: StepGraph records them while it warms the step up and captures only once they repeat
scalarTrace = None


def traceScalar(tag, *values):
	if scalarTrace is not None:
		scalarTrace.append((tag, ) + tuple(float(v) for v in values))


_graphsInFlight = []


def syncGraphs():
	"""host reads / writes of device memory go through the legacy stream, which does not wait for the (non-blocking) stream a
	StepGraph replays on: synchronise every graph with replays in flight first (GPUArray.get / set call this)"""
	while _graphsInFlight:
		_graphsInFlight.pop().stream.synchronize()


currentStream = None          # the Stream every operator call runs on (None: the legacy default stream)

# hook of the data-parallel gradient synchronisation (grid.GradientSync.noteWrite): called with every array a backend op is
# about to write parameter gradients into; None outside a grid
gradientWriteHook = None


def setDefaultStream(stream):
	"""Route every operator call that does not name a stream (all of the reference API) to `stream`; None restores the legacy
	default stream."""
	global currentStream
	flushDeferred()
	currentStream = stream
	check(lib.pz_set_default_stream(stream.handle if stream is not None else None))


class StepGraph:
	"""One training / inference step captured as a CUDA graph.

	`fn` is any callable driving the unchanged operator API (e.g. zeroGradParams + net(data) + net.backward(grad) +
	optimizer.update() + net.reset()).  It is run eagerly on a private stream at least `warmup` times -- so that the memory pool
	holds every block the step needs and the library scratch has its final size -- and further until the scalars that the step
	hands to the kernels (batch-norm running-average factor max(1/n, minFactor), optimizer rates) repeat from one run to the
	next, at most `maxWarmup` times; a step whose scalars never settle (Adam's bias-corrected rate) cannot be replayed and raises.
	Then it is run once more under stream capture.  `launch()` replays the captured kernels with a single host call: device
	pointers, shapes and scalars are those of the captured run; the dropout generator keeps its offset on the device, so replays
	draw fresh masks.  `GPUArray.get()` / `set()` wait for replays in flight; other host-side synchronisation is the caller's
	(`synchronize()`).
	"""

	def __init__(self, fn, warmup=2, maxWarmup=64):
		global scalarTrace
		self.stream = Stream()
		self.exec = None
		setDefaultStream(self.stream)
		try:
			previous, runs = None, 0
			while True:
				scalarTrace = []
				try:
					fn()
				finally:
					trace, scalarTrace = scalarTrace, None
				runs += 1
				if runs >= warmup and trace == previous:
					break
				if runs >= maxWarmup:
					changing = sorted({a[0] for a, b in zip(trace, previous or []) if a != b}) or ["the launch sequence"]
					raise RuntimeError("the step cannot be captured: %s still change(s) from step to step after %d warm-up runs"
									   % (", ".join(changing), runs))
				previous = trace
			self.stream.synchronize()
			check(lib.pz_graph_begin(self.stream.handle))
			try:
				fn()
				flushDeferred()
			finally:
				h = c_void_p()
				status = lib.pz_graph_end(self.stream.handle, byref(h))
			check(status)
			self.exec = h.value
			self.warmupRuns = runs
		finally:
			setDefaultStream(None)

	def launch(self):
		flushDeferred()
		check(lib.pz_graph_launch(self.exec, self.stream.handle))
		if self not in _graphsInFlight:
			_graphsInFlight.append(self)

	def synchronize(self):
		self.stream.synchronize()
		if self in _graphsInFlight:
			_graphsInFlight.remove(self)

	def destroy(self):
		"""Release the graph (graphs that captured NCCL collectives must be destroyed BEFORE their communicator)."""
		if self.exec:
			self.synchronize()
			lib.pz_graph_destroy(self.exec)
			self.exec = None

	def __del__(self):
		try:
			self.destroy()
		except Exception:
			pass


class Event:
	def __init__(self, timing=True):
		h = c_void_p()
		check(lib.pz_event_create(byref(h)) if timing else lib.pz_event_create_sync(byref(h)))
		self.handle = h.value

	def record(self, stream=None):
		flushDeferred()
		check(lib.pz_event_record(self.handle, stream.handle if stream is not None else None))

	def synchronize(self):
		check(lib.pz_event_synchronize(self.handle))

	def timeTill(self, other):
		ms = c_float(0)
		check(lib.pz_event_elapsed_ms(self.handle, other.handle, byref(ms)))
		return ms.value

	timeSince = lambda self, other: other.timeTill(self)

	def __del__(self):
		try:
			if self.handle:
				lib.pz_event_destroy(self.handle)
				self.handle = None
		except Exception:
			pass
