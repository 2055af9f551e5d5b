"""GPUArray -- n-d device array with numpy-style shape / byte strides over a driver.Buffer.

Mirrors the reference type `Cuda.Driver.GPUArray` (Cuda/Source/Core/Array.c, Driver.h:280-292) together with the
Python-side extension of Cuda/GPUArray.py (fill / astype / min / max / + * += *=), implemented on the C-ABI of
libpzb200.so.  Views share the Buffer; non-contiguous views (results of slicing) can be read, written and copied
through pitched 2-D copies exactly like the reference does with memcpy2D/3D (Array.c:560-801).
"""
import os

import numpy as np

from . import driver
from .driver import lib, check, dtypeCode, Buffer

_f32 = np.dtype(np.float32)
_memcpy_d2d = lib.pz_memcpy_d2d


def _prod(shape):
	n = 1
	for d in shape:
		n *= d
	return n


def _cstrides(shape, itemsize):
	strides = [0] * len(shape)
	acc = itemsize
	for i in range(len(shape) - 1, -1, -1):
		strides[i] = acc
		acc *= shape[i]
	return tuple(strides)


DEFER_MIN_BYTES = 1 << 20


class DeferredFill:
	"""The one launch the backend may hold back (see driver.flushDeferred): `y.fill(0)`, which `accumulate` turns into the scaled
	copy `y = 0 + a * x` and then into the fused `y = (0 + a1 * x1) + a2 * x2`.  Holds references so that neither array's memory
	can return to the pool while the launch is pending."""
	__slots__ = ["y", "x", "alpha", "x2", "alpha2"]

	def __init__(self, y):
		self.y, self.x, self.alpha, self.x2, self.alpha2 = y, None, 0.0, None, 0.0

	def matches(self, y, x):
		return self.y._ptr == y._ptr and self.y.nbytes == y.nbytes and self.y.dtype == y.dtype and x.dtype == y.dtype and \
			x.size == y.size and x.contiguous and y.contiguous and (x._ptr + x.nbytes <= y._ptr or y._ptr + y.nbytes <= x._ptr)

	def flush(self):
		y = self.y
		if self.x is None:
			check(lib.pz_memset8(y._ptr, 0, y.nbytes, None))
		elif self.x2 is None:
			check(lib.pz_axpy2(dtypeCode(y.dtype), y._ptr, self.x._ptr, self.alpha, None, 0.0, y.size, None))
		else:
			check(lib.pz_axpy2(dtypeCode(y.dtype), y._ptr, self.x._ptr, self.alpha, self.x2._ptr, self.alpha2, y.size, None))


class DeferredBatchNorm:
	"""A train-mode batch-norm forward launch held back for one call: when the next operator is the ReLU of its output (conv -> bn ->
	relu inside every ResNet block) one pass stores both (`pz_bn_fwd_train_relu`); any other access to device memory launches it as
	it is.  Holds references to every tensor of the launch."""
	__slots__ = ["code", "x", "y", "geometry", "params", "eps", "factor"]

	def __init__(self, code, x, y, geometry, params, eps, factor):
		self.code, self.x, self.y, self.geometry, self.params, self.eps, self.factor = code, x, y, geometry, params, eps, factor

	def launch(self, z=None):
		ptrs = [p._ptr for p in self.params]          # scale, bias, mean, var, savemean, saveinvvar
		if z is None:
			check(lib.pz_bn_fwd_train(self.code, self.x._ptr, self.y._ptr, *self.geometry, *ptrs, self.eps, self.factor, None))
		else:
			check(lib.pz_bn_fwd_train_relu(self.code, self.x._ptr, self.y._ptr, z._ptr, *self.geometry, *ptrs, self.eps, self.factor, None))

	def flush(self):
		self.launch()


class DeferredBatchNormBackward:
	"""A batch-norm backward launch held back for up to two calls: BatchNormND.accGradParams (Modules/BatchNormND.py:74-83) follows it
	with `addVectorToVector(scalegrad, acc, out=acc, alpha, beta)` and the same for the bias gradient -- two launches over C floats,
	106 of them in a ResNet-50 step.  `accumulateAfterBatchNormBackward` absorbs them and the pass writes the accumulators itself
	(`pz_bn_bwd_acc`, same bits); any other access to device memory launches the pass as it is."""
	__slots__ = ["code", "x", "dy", "dx", "geometry", "params", "scalegrad", "bgrad", "sacc", "bacc"]

	def __init__(self, code, x, dy, dx, geometry, params, scalegrad, bgrad):
		self.code, self.x, self.dy, self.dx, self.geometry, self.params = code, x, dy, dx, geometry, params
		self.scalegrad, self.bgrad, self.sacc, self.bacc = scalegrad, bgrad, None, None

	def tensors(self):
		return (self.x, self.dy, self.dx, self.scalegrad, self.bgrad) + tuple(self.params) + \
			tuple(acc[0] for acc in (self.sacc, self.bacc) if acc is not None)

	def flush(self):
		ptrs = [p._ptr for p in self.params]          # scale, savemean, saveinvvar
		head = (self.code, self.x._ptr, self.dy._ptr, self.dx._ptr, *self.geometry, *ptrs, self.scalegrad._ptr, self.bgrad._ptr)
		if self.sacc is None and self.bacc is None:
			check(lib.pz_bn_bwd(*head, None))
			return
		acc = []
		for a in (self.sacc, self.bacc):
			acc += [None, 0.0, 0.0] if a is None else [a[0]._ptr, a[1], a[2]]
		check(lib.pz_bn_bwd_acc(*head, *acc, None))


_NO_BN_ACC = bool(int(os.environ.get("PZ_NO_BN_ACC_FUSION", "0")))      # A/B switch for timing and for the bit-identity test


def accumulateAfterBatchNormBackward(out, x, alpha, y, beta):
	"""out = alpha * x + beta * y (the reference's addKer) where x is a parameter gradient of the pending batch-norm backward pass
	and out is y, a float32 accumulator apart from every tensor of the pass: absorbed; True when it was"""
	op = driver.deferred
	if _NO_BN_ACC or type(op) is not DeferredBatchNormBackward or out._ptr != y._ptr or out.nbytes != y.nbytes or out.dtype != np.float32 or \
			y.dtype != np.float32 or x.dtype != np.float32 or out.size != x.size or not (out.contiguous and x.contiguous and y.contiguous):
		return False
	if op.sacc is None and x._ptr == op.scalegrad._ptr and x.nbytes == op.scalegrad.nbytes:
		slot = "sacc"
	elif op.bacc is None and x._ptr == op.bgrad._ptr and x.nbytes == op.bgrad.nbytes:
		slot = "bacc"
	else:
		return False
	for other in op.tensors():
		if not (out._ptr + out.nbytes <= other._ptr or other._ptr + other.nbytes <= out._ptr):
			return False
	setattr(op, slot, (out, float(alpha), float(beta)))
	if op.sacc is not None and op.bacc is not None:
		driver.flushDeferred()                        # nothing more to wait for
	return True


def reluAfterBatchNorm(out, inp):
	"""relu(out, inp) where `inp` is the output of a pending batch-norm launch: fused; True when it did"""
	op = driver.deferred
	if _NO_SUM_RELU or type(op) is not DeferredBatchNorm or op.y._ptr != inp._ptr or op.y.nbytes != inp.nbytes or out.dtype != inp.dtype or \
			out.size != inp.size or not out.contiguous:
		return False
	for other in (op.x, op.y) + tuple(op.params):
		if not (out._ptr + out.nbytes <= other._ptr or other._ptr + other.nbytes <= out._ptr):
			return False
	driver.deferred = None
	op.launch(out)
	return True


def accumulate(y, x, alpha):
	"""y += alpha * x (the reference's toVectorAddVector kernel) with the pending-fill fusion; returns True when the launch was
	absorbed or issued here"""
	op = driver.deferred
	if type(op) is not DeferredFill or not op.matches(y, x):
		return False
	if op.x is None:
		op.x, op.alpha = x, float(alpha)                      # y = 0 + alpha * x, still pending
		return True
	if op.x2 is None:
		# y = (0 + a1 * x1) + a2 * x2: complete for a two-input Add / Replicate; kept pending one more call -- if the next
		# operator is the ReLU of this sum (Add -> Activation in every ResNet block), `reluAfterSum` stores both in one pass
		op.x2, op.alpha2 = x, float(alpha)
		return True
	driver.flushDeferred()                                    # a third term: the two-term sum goes out, this one takes the plain kernel
	return False


_NO_SUM_RELU = bool(int(os.environ.get("PZ_NO_SUM_RELU_FUSION", "0")))      # A/B switch for timing


def _pendingSum(inp, out, *others):
	"""the pending two-term sum that `inp` is, if `out` (a distinct tensor of the same layout) can be written next to it"""
	op = driver.deferred
	if _NO_SUM_RELU or type(op) is not DeferredFill or op.x2 is None or op.y._ptr != inp._ptr or op.y.nbytes != inp.nbytes or out.dtype != inp.dtype or \
			out.size != inp.size or not out.contiguous:
		return None
	for other in (inp, op.x, op.x2) + others:
		if not (out._ptr + out.nbytes <= other._ptr or other._ptr + other.nbytes <= out._ptr):
			return None
	return op


def reluAfterSum(out, inp):
	"""relu(out, inp) where `inp` is a pending two-term sum: one launch writes the sum and its ReLU; True when it did"""
	op = _pendingSum(inp, out)
	if op is None:
		return False
	driver.deferred = None
	check(lib.pz_axpy2_relu(dtypeCode(inp.dtype), inp._ptr, out._ptr, op.x._ptr, op.alpha, op.x2._ptr, op.alpha2, inp.size, None))
	return True


def reluDerAfterSum(ingrad, outgrad, ref):
	"""reluDer(ingrad, outgrad, ref) where `outgrad` is a pending two-term sum (Replicate's gradient sum feeding the previous block's
	ReLU): one launch writes the sum and ingrad = sum * (ref > 0); True when it did"""
	if ref.dtype != outgrad.dtype or ref.size != outgrad.size or not ref.contiguous:
		return False
	op = _pendingSum(outgrad, ingrad, ref)
	if op is None or not (ref._ptr + ref.nbytes <= outgrad._ptr or outgrad._ptr + outgrad.nbytes <= ref._ptr):
		return False
	driver.deferred = None
	check(lib.pz_axpy2_relu_bwd(dtypeCode(outgrad.dtype), outgrad._ptr, ingrad._ptr, op.x._ptr, op.alpha, op.x2._ptr, op.alpha2, ref._ptr,
								outgrad.size, None))
	return True


class GPUArray:
	__slots__ = ["shape", "strides", "dtype", "gpudata", "_ptr", "size", "contiguous", "__weakref__"]

	MAXDIMS = 32  # reference: Driver.h:234-237

	def __init__(self, shape, dtype, allocator=None, gpudata=None, strides=None, offset=0):
		if isinstance(shape, (int, np.integer)):
			shape = (int(shape), )
		else:
			shape = tuple(int(d) for d in shape)

		if len(shape) > self.MAXDIMS:
			raise ValueError("invalid number of dimensions")

		dtype = dtype if type(dtype) is np.dtype else np.dtype(dtype)

		self.shape = shape
		self.dtype = dtype
		self.size = _prod(shape)

		if strides is None:
			self.strides = _cstrides(shape, dtype.itemsize)
			self.contiguous = True
		else:
			self.strides = tuple(strides)
			self.contiguous = self.strides == _cstrides(shape, dtype.itemsize) or self.size <= 1

		if gpudata is None:
			gpudata = Buffer(self.size * dtype.itemsize, allocator=allocator)
		elif self.contiguous and gpudata.size - offset < self.size * dtype.itemsize:
			raise ValueError("gpudata is too small for the requested array")

		self.gpudata = gpudata
		self._ptr = gpudata.ptr + offset

	# ------------------------------------------------------------------------------------------ constructors
	@staticmethod
	def empty(shape, dtype, allocator=None, gpudata=None):
		return GPUArray(shape, dtype, allocator=allocator, gpudata=gpudata)

	@staticmethod
	def zeros(shape, dtype, allocator=None, gpudata=None):
		ary = GPUArray(shape, dtype, allocator=allocator, gpudata=gpudata)
		if ary.size > 0:
			check(lib.pz_memset8(ary.ptr, 0, ary.nbytes, None))
		return ary

	@staticmethod
	def emptyLike(ary, allocator=None):
		return GPUArray(ary.shape, ary.dtype, allocator=allocator)

	@staticmethod
	def zerosLike(ary, allocator=None):
		return GPUArray.zeros(ary.shape, ary.dtype, allocator=allocator)

	@staticmethod
	def toGpu(host, allocator=None):
		host = np.ascontiguousarray(host)
		ary = GPUArray(host.shape, host.dtype, allocator=allocator)
		if ary.size > 0:
			check(lib.pz_memcpy_h2d(ary.ptr, host.ctypes.data, host.nbytes, None, 0))
		return ary

	# ------------------------------------------------------------------------------------------ properties
	@property
	def ptr(self):
		"""device address of the first element.  Handing it out means somebody is about to touch device memory: a pending
		deferred fill (driver.flushDeferred) is issued first"""
		if driver.deferred is not None:
			driver.flushDeferred()
		return self._ptr

	@property
	def ndim(self):
		return len(self.shape)

	@property
	def nbytes(self):
		return self.size * self.dtype.itemsize

	@property
	def itemsize(self):
		return self.dtype.itemsize

	@property
	def device(self):
		return driver.Device.getCurrent()

	@property
	def parent(self):
		return self.gpudata

	def dimAt(self, index):
		return self.shape[index]

	def strideAt(self, index):
		return self.strides[index]

	def __len__(self):
		if len(self.shape) == 0:
			raise TypeError("len() of a 0-d gpuarray")
		return self.shape[0]

	def __repr__(self):
		return "GPUArray(shape=%s, dtype=%s, contiguous=%s)" % (self.shape, self.dtype, self.contiguous)

	# ------------------------------------------------------------------------------------------ host transfer
	def _chunks(self):
		"""Decompose a strided view into (offset, rows, rowbytes, pitch) pitched blocks."""
		itemsize = self.dtype.itemsize
		shape, strides = [], []
		for d, s in zip(self.shape, self.strides):
			if d != 1:
				shape.append(d)
				strides.append(s)

		# merge dims that are contiguous with their inner neighbour
		mshape, mstrides = [], []
		for d, s in zip(reversed(shape), reversed(strides)):
			if mshape and s == mshape[-1] * mstrides[-1]:
				mshape[-1] *= d
			else:
				mshape.append(d)
				mstrides.append(s)

		if not mshape:
			return [(0, 1, itemsize, itemsize)]

		if mstrides[0] != itemsize:
			mshape.insert(0, 1)
			mstrides.insert(0, itemsize)

		rowbytes = mshape[0] * itemsize
		if len(mshape) == 1:
			return [(0, 1, rowbytes, rowbytes)]

		rows, pitch = mshape[1], mstrides[1]
		outer = list(zip(mshape[2:], mstrides[2:]))

		offsets = [0]
		for d, s in outer:
			offsets = [o + i * s for i in range(d) for o in offsets]

		return [(o, rows, rowbytes, pitch) for o in sorted(offsets)]

	def get(self, stream=None):
		host = np.empty(self.shape, dtype=self.dtype)
		if self.size == 0:
			return host
		driver.syncGraphs()

		if self.contiguous:
			check(lib.pz_memcpy_d2h(host.ctypes.data, self.ptr, self.nbytes, None, 0))
			return host

		hostptr = host.ctypes.data
		for offset, rows, rowbytes, pitch in self._chunks():
			check(lib.pz_memcpy2d(hostptr, rowbytes, self.ptr + offset, pitch, rowbytes, rows, 2, None))
			hostptr += rows * rowbytes
		return host

	def set(self, ary, stream=None):
		if isinstance(ary, GPUArray):
			if ary.shape != self.shape or ary.dtype != self.dtype:
				raise ValueError("gpuarray shapes or datatypes are not equal")
			ary.copy(out=self)
			return self

		host = np.ascontiguousarray(ary, dtype=self.dtype)
		if host.shape != self.shape:
			if host.size != self.size:
				raise ValueError("host array has invalid shape %s (expected %s)" % (host.shape, self.shape))
			host = host.reshape(self.shape)

		if self.size == 0:
			return self
		driver.syncGraphs()

		if self.contiguous:
			check(lib.pz_memcpy_h2d(self.ptr, host.ctypes.data, host.nbytes, None, 0))
			return self

		hostptr = host.ctypes.data
		for offset, rows, rowbytes, pitch in self._chunks():
			check(lib.pz_memcpy2d(self.ptr + offset, pitch, hostptr, rowbytes, rowbytes, rows, 1, None))
			hostptr += rows * rowbytes
		return self

	def copy(self, allocator=None, out=None):
		"""Contiguous copy of this array (or copy into `out`, which may itself be a strided view)."""
		if out is None:
			out = GPUArray(self.shape, self.dtype, allocator=allocator)

		if self.size == 0:
			return out

		if self.contiguous and out.contiguous:
			check(_memcpy_d2d(out.ptr, self.ptr, self.nbytes, None))
			return out

		if self.contiguous or out.contiguous:
			strided, dense = (out, self) if self.contiguous else (self, out)
			denseptr = dense.ptr
			for offset, rows, rowbytes, pitch in strided._chunks():
				if strided is out:
					check(lib.pz_memcpy2d(out.ptr + offset, pitch, denseptr, rowbytes, rowbytes, rows, 0, None))
				else:
					check(lib.pz_memcpy2d(denseptr, rowbytes, self.ptr + offset, pitch, rowbytes, rows, 0, None))
				denseptr += rows * rowbytes
			return out

		tmp = self.copy(allocator=allocator)
		return tmp.copy(out=out)

	# ------------------------------------------------------------------------------------------ views
	def reshape(self, *shape):
		if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
			shape = tuple(shape[0])
		shape = [int(d) for d in shape]

		unknown = [i for i, d in enumerate(shape) if d == -1]
		if len(unknown) > 1:
			raise ValueError("only one dimension can be inferred")
		if unknown:
			known = _prod(d for d in shape if d != -1)
			if known == 0 or self.size % known != 0:
				raise ValueError("cannot reshape gpuarray of size %d into shape %s" % (self.size, tuple(shape)))
			shape[unknown[0]] = self.size // known

		shape = tuple(shape)
		if _prod(shape) != self.size:
			raise ValueError("total size of new gpuarray must be unchanged")

		if self.contiguous:
			return GPUArray(shape, self.dtype, gpudata=self.gpudata, offset=self._ptr - self.gpudata.ptr)

		strides = self._reshapeStrides(shape)
		if strides is None:
			raise ValueError("cannot reshape non-contiguous gpuarray without a copy")
		return GPUArray(shape, self.dtype, gpudata=self.gpudata, strides=strides, offset=self._ptr - self.gpudata.ptr)

	def _reshapeStrides(self, newshape):
		# numpy's no-copy reshape rule
		olddims = [(d, s) for d, s in zip(self.shape, self.strides) if d != 1]
		newstrides = [0] * len(newshape)
		oi, ni = 0, 0
		on, nn = len(olddims), len(newshape)

		while oi < on and ni < nn:
			np_, op_ = newshape[ni], olddims[oi][0]
			oj, nj = oi + 1, ni + 1
			while np_ != op_:
				if np_ < op_:
					np_ *= newshape[nj]
					nj += 1
				else:
					op_ *= olddims[oj][0]
					oj += 1
			for k in range(oi, oj - 1):
				if olddims[k][1] != olddims[k + 1][0] * olddims[k + 1][1]:
					return None
			newstrides[nj - 1] = olddims[oj - 1][1]
			for k in range(nj - 1, ni, -1):
				newstrides[k - 1] = newstrides[k] * newshape[k]
			oi, ni = oj, nj

		last = newstrides[ni - 1] if ni > 0 else self.dtype.itemsize
		for k in range(ni, nn):
			newstrides[k] = last
		return tuple(newstrides)

	def ravel(self):
		return self.reshape(self.size)

	def view(self, dtype):
		dtype = np.dtype(dtype)
		if not self.contiguous:
			raise ValueError("gpuarray is not contiguous")
		if dtype.itemsize == self.dtype.itemsize:
			shape = self.shape
		else:
			lastbytes = self.shape[-1] * self.dtype.itemsize if self.shape else self.dtype.itemsize
			if lastbytes % dtype.itemsize != 0:
				raise ValueError("last axis size is not divisible by the new itemsize")
			shape = self.shape[:-1] + (lastbytes // dtype.itemsize, )
		return GPUArray(shape, dtype, gpudata=self.gpudata, offset=self._ptr - self.gpudata.ptr)

	def __getitem__(self, key):
		if not isinstance(key, tuple):
			key = (key, )

		if any(k is Ellipsis for k in key):
			idx = next(i for i, k in enumerate(key) if k is Ellipsis)
			nspec = sum(1 for k in key if k is not None and k is not Ellipsis)
			key = key[:idx] + (slice(None), ) * (len(self.shape) - nspec) + key[idx + 1:]

		shape, strides, offset = [], [], 0
		axis = 0

		for k in key:
			if k is None:
				shape.append(1)
				strides.append(strides[-1] if strides else (self.strides[axis] * self.shape[axis] if axis < len(self.shape)
															 else self.dtype.itemsize))
				continue

			if axis >= len(self.shape):
				raise IndexError("too many indices for gpuarray")

			dim, stride = self.shape[axis], self.strides[axis]

			if isinstance(k, (int, np.integer)):
				k = int(k)
				if k < 0:
					k += dim
				if not 0 <= k < dim:
					raise IndexError("index out of range")
				offset += k * stride

			elif isinstance(k, slice):
				start, stop, step = k.indices(dim)
				if step != 1:
					raise ValueError("slice step is not supported")
				shape.append(max(0, stop - start))
				strides.append(stride)
				offset += start * stride

			else:
				raise TypeError("invalid index type %s" % type(k).__name__)

			axis += 1

		shape.extend(self.shape[axis:])
		strides.extend(self.strides[axis:])

		return GPUArray(shape, self.dtype, gpudata=self.gpudata, strides=strides, offset=self._ptr - self.gpudata.ptr + offset)

	def __setitem__(self, key, value):
		self[key].set(value)

	# ------------------------------------------------------------------------------------------ arithmetic
	def enforceContiguous(self):
		if not self.contiguous:
			raise ValueError("gpuarray is not contiguous")

	def _findAllocator(self, *others):
		for ary in (self, ) + others:
			buf = ary.gpudata
			while buf is not None:
				if buf.pool is not None:
					return buf.pool
				buf = buf.parent
		return None

	def _enforceEqual(self, other):
		self.enforceContiguous()
		other.enforceContiguous()
		if self.shape != other.shape:
			raise ValueError("gpuarray shapes are not equal")
		if self.dtype != other.dtype:
			raise ValueError("gpuarray datatypes are not equal")

	def fill(self, val):
		# raw bit-pattern memset for 1/2/4-byte types, kernel for 8-byte ones (reference: Cuda/GPUArray.py:167-180)
		self.enforceContiguous()
		if self.size == 0:
			return self

		item = np.array(val).astype(self.dtype)
		itemsize = self.dtype.itemsize

		if self.nbytes >= DEFER_MIN_BYTES and itemsize in (2, 4) and not item.tobytes().strip(b"\0"):
			# a zero fill of a large tensor: keep it pending, the accumulation that follows usually absorbs it
			driver.flushDeferred()
			driver.deferred = DeferredFill(self)
			return self

		if itemsize == 4:
			check(lib.pz_memset32(self.ptr, int(item.view(np.uint32)), self.size, None))
		elif itemsize == 2:
			check(lib.pz_memset16(self.ptr, int(item.view(np.uint16)), self.size, None))
		elif itemsize == 1:
			check(lib.pz_memset8(self.ptr, int(item.view(np.uint8)), self.size, None))
		elif itemsize == 8:
			check(lib.pz_fill64(self.ptr, int(item.view(np.uint64)), self.size, None))
		else:
			raise NotImplementedError("fill for itemsize %d" % itemsize)
		return self

	def astype(self, dtype):
		self.enforceContiguous()
		dtype = np.dtype(dtype)
		out = GPUArray(self.shape, dtype, allocator=self._findAllocator())
		if self.dtype == dtype:
			out.set(self)
		else:
			check(lib.pz_cast(dtypeCode(dtype), out.ptr, dtypeCode(self.dtype), self.ptr, self.size, None))
		return out

	def _reduce(self, wantMax):
		self.enforceContiguous()
		out = GPUArray((), self.dtype, allocator=self._findAllocator())
		if self.dtype == np.int32:
			check(lib.pz_reduce_minmax_i32(self.ptr, self.size, 1 if wantMax else 0, out.ptr, None))
		else:
			check(lib.pz_reduce_minmax(dtypeCode(self.dtype), self.ptr, self.size, 1 if wantMax else 0, out.ptr, None))
		return out

	def min(self):
		return self._reduce(False)

	def max(self):
		return self._reduce(True)

	def __add__(self, other):
		self._enforceEqual(other)
		out = GPUArray(self.shape, self.dtype, allocator=self._findAllocator(other))
		check(lib.pz_axpby(dtypeCode(self.dtype), out.ptr, self.ptr, 1.0, other.ptr, 1.0, self.size, None))
		return out

	def __mul__(self, other):
		self._enforceEqual(other)
		out = GPUArray(self.shape, self.dtype, allocator=self._findAllocator(other))
		check(lib.pz_mul(dtypeCode(self.dtype), out.ptr, self.ptr, other.ptr, self.size, None))
		return out

	def __iadd__(self, other):
		self._enforceEqual(other)
		check(lib.pz_axpy(dtypeCode(self.dtype), self.ptr, other.ptr, 1.0, self.size, None))
		return self

	def __imul__(self, other):
		self._enforceEqual(other)
		check(lib.pz_mul(dtypeCode(self.dtype), self.ptr, self.ptr, other.ptr, self.size, None))
		return self
