"""Data-parallel gradient synchronisation: the `NodeInfo` of the reference's Grid.py re-done over NCCL.

The reference runs one process per GPU and averages the flat gradient buffer through a parent/child STAR over CUDA-IPC
mapped memory with `multiprocessing.SimpleQueue`s as the control plane (reference: Grid.py:4-157).  Here the process
model and the NodeInfo surface stay -- `sumTensor(name, tensor)`, `broadcastBuffer(name, buffer)`, `meanValue(value)`,
`index`, `gridsize`, `device`, `close()` -- so `Optimizer` is untouched, while the data plane is ONE ncclAllReduce /
ncclBroadcast per call over NVLink / NVSwitch and the control plane (exchange of the 128-byte ncclUniqueId, scalar
means, barriers) is a `Rendezvous` object: `SocketRendezvous` (plain TCP, what `runGrid` uses -- like the reference's
queues it needs nothing beyond the standard library) or `TorchRendezvous` over torch.distributed's gloo backend for processes
that `torchrun` started (bench.py).  No tensor of the hot path ever enters either.
"""
import ctypes
import os
from ctypes import byref

from . import driver
from .driver import lib, check, dtypeCode


class SocketRendezvous:
	"""Host-side control plane over plain TCP: a star through rank 0 (the role of the parent's SimpleQueues in Grid.py:14-57).

	Every collective is one round trip: the other ranks send their contribution to rank 0, which combines and answers."""

	def __init__(self, rank, size, masterAddr="127.0.0.1", masterPort=29533, timeout=600):
		import socket
		import time

		self.rank, self.size = int(rank), int(size)
		self.peers = {}

		if self.rank == 0:
			server = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
			server.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
			server.bind((masterAddr, int(masterPort)))
			server.listen(self.size)
			server.settimeout(timeout)
			try:
				while len(self.peers) < self.size - 1:
					conn, _ = server.accept()
					conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
					conn.settimeout(timeout)
					peer = self._recv(conn)
					if not isinstance(peer, int) or not 0 < peer < self.size or peer in self.peers:
						raise RuntimeError("rendezvous: unexpected hello %r" % (peer, ))
					self.peers[peer] = conn
			finally:
				server.close()
		else:
			deadline = time.time() + timeout
			while True:
				conn = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
				try:
					conn.connect((masterAddr, int(masterPort)))
					break
				except OSError:
					conn.close()
					if time.time() > deadline:
						raise
					time.sleep(0.05)
			conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
			conn.settimeout(timeout)
			self._send(conn, self.rank)
			self.peers[0] = conn

	@staticmethod
	def _send(conn, obj):
		import pickle
		import struct
		payload = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
		conn.sendall(struct.pack("<Q", len(payload)) + payload)

	@staticmethod
	def _recv(conn):
		import pickle
		import struct

		def exactly(n):
			chunks = []
			while n > 0:
				chunk = conn.recv(n)
				if not chunk:
					raise ConnectionError("rendezvous: peer closed the connection")
				chunks.append(chunk)
				n -= len(chunk)
			return b"".join(chunks)

		(n, ) = struct.unpack("<Q", exactly(8))
		return pickle.loads(exactly(n))

	def _collective(self, value, combine):
		"""every rank contributes `value`; all get combine([values in rank order])"""
		if self.size == 1:
			return combine([value])
		if self.rank == 0:
			values = [value] + [self._recv(self.peers[r]) for r in range(1, self.size)]
			result = combine(values)
			for r in range(1, self.size):
				self._send(self.peers[r], result)
			return result
		self._send(self.peers[0], value)
		return self._recv(self.peers[0])

	def broadcastBytes(self, payload, root=0):
		return self._collective(payload if self.rank == root else None, lambda values: values[root])

	def barrier(self):
		self._collective(None, lambda values: None)

	def maxValue(self, value):
		return self._collective(float(value), max)

	def sumValue(self, value):
		return self._collective(float(value), sum)

	def meanValue(self, value):
		return self.sumValue(value) / self.size

	def close(self):
		for conn in self.peers.values():
			try:
				conn.close()
			except OSError:
				pass
		self.peers = {}


class TorchRendezvous:
	"""The same control plane over torch.distributed (gloo), for processes that torchrun launched and initialised."""

	def __init__(self, rank=None, size=None, masterAddr=None, masterPort=None, timeout=600):
		import datetime
		import torch.distributed as dist

		self.dist = dist
		self.owned = False

		if not dist.is_initialized():
			if rank is not None:
				os.environ["RANK"], os.environ["WORLD_SIZE"] = str(rank), str(size)
				os.environ["MASTER_ADDR"] = masterAddr or os.environ.get("MASTER_ADDR", "127.0.0.1")
				os.environ["MASTER_PORT"] = str(masterPort or os.environ.get("MASTER_PORT", "29533"))
			dist.init_process_group("gloo", timeout=datetime.timedelta(seconds=timeout))
			self.owned = True

		self.rank, self.size = dist.get_rank(), dist.get_world_size()

	def broadcastBytes(self, payload, root=0):
		box = [payload if self.rank == root else None]
		self.dist.broadcast_object_list(box, src=root)
		return box[0]

	def barrier(self):
		self.dist.barrier()

	def _reduce(self, value, op):
		import torch
		t = torch.tensor([float(value)], dtype=torch.float64)
		self.dist.all_reduce(t, op=op)
		return float(t[0])

	def maxValue(self, value):
		return self._reduce(value, self.dist.ReduceOp.MAX)

	def sumValue(self, value):
		return self._reduce(value, self.dist.ReduceOp.SUM)

	def meanValue(self, value):
		return self.sumValue(value) / self.size

	def close(self):
		if self.owned and self.dist.is_initialized():
			self.dist.destroy_process_group()
			self.owned = False


class NcclCommunicator:
	"""One NCCL communicator per process, created from a unique id that rank 0 publishes through the rendezvous."""

	def __init__(self, rendezvous):
		self.rank, self.size = rendezvous.rank, rendezvous.size

		uid = ctypes.create_string_buffer(128)
		if self.rank == 0:
			check(lib.pz_nccl_unique_id(uid))
		payload = rendezvous.broadcastBytes(uid.raw if self.rank == 0 else None, root=0)
		uid = ctypes.create_string_buffer(payload, 128)

		handle = ctypes.c_void_p()
		check(lib.pz_nccl_comm_init(byref(handle), self.size, self.rank, uid))
		self.handle = handle.value

	def allReduceMean(self, ary):
		check(lib.pz_nccl_allreduce_mean(self.handle, dtypeCode(ary.dtype), ary.ptr, ary.size, 1.0 / self.size, None))

	def broadcastBytes(self, ptr, nbytes, root=0):
		check(lib.pz_nccl_broadcast(self.handle, driver.PZ_U8, ptr, nbytes, root, None))

	def allReduceMomentumSGD(self, param, grad, mom, learnRate, momRate):
		check(lib.pz_nccl_allreduce_sgd_momentum(self.handle, dtypeCode(param.dtype), param.ptr, grad.ptr, mom.ptr, param.size,
												 1.0 / self.size, learnRate, momRate, None))

	def close(self):
		if self.handle:
			lib.pz_nccl_comm_destroy(self.handle)
			self.handle = None


class GradientSync:
	"""Overlaps the gradient mean with the backward pass (the reference reduces after backward, Grid.py:123-157).

	The flat gradient buffer of `Optimizer.setupGlobalState` (Optimizers/Optimizer.py:66-111) is learned from the first
	`sumTensor` call.  From then on every backend op that writes parameter gradients (convNdBackwardParams, gemm / matsum /
	addKer with an `out` inside the buffer) reports the range it is about to write; once `bucketBytes` of freshly written
	ranges have piled up they are averaged over the ranks by ONE grouped NCCL launch on a communication stream, ordered after
	the kernels that wrote them by an event, while the compute stream goes on with the earlier layers.  `finish` (the
	reference's `sumTensor` call in `Optimizer.updateGlobalState`) reduces whatever is not known to be clean and makes the
	compute stream wait for the communication stream.

	Safety: a range that is written again after it was reduced (accumulation into a shared variable) is simply dirty again --
	the mean of (mean(g1) + g2_r) over the ranks is mean(g1) + mean(g2), so reducing twice is exact; the writer first waits for
	a reduction of that range still in flight.  Ranges nobody reported are never assumed clean.  The kernels and the NCCL
	launches of a step are the same on every rank in the same order (same model, same shapes), as NCCL requires."""

	def __init__(self, comm, bucketBytes=24 << 20):
		self.comm = comm
		self.bucketBytes = bucketBytes
		self.regions = {}                # ptr of the flat buffer -> [ptr, nbytes, dtype, clean ranges, dirty ranges, dirty bytes]
		self.stream = driver.Stream()
		self.inflight = []               # (event, [(lo, hi)]) of buckets whose reduction may still be running
		self.launches = 0

	# ---- range bookkeeping (byte addresses, half-open)
	@staticmethod
	def _merge(ranges):
		out = []
		for lo, hi in sorted(ranges):
			if out and lo <= out[-1][1]:
				out[-1][1] = max(out[-1][1], hi)
			else:
				out.append([lo, hi])
		return out

	@staticmethod
	def _subtract(ranges, cut):
		lo, hi = cut
		out = []
		for a, b in ranges:
			if b <= lo or a >= hi:
				out.append([a, b])
			else:
				if a < lo:
					out.append([a, lo])
				if b > hi:
					out.append([hi, b])
		return out

	def _region(self, ptr):
		for reg in self.regions.values():
			if reg[0] <= ptr < reg[0] + reg[1]:
				return reg
		return None

	def noteWrite(self, ary):
		"""`ary` (a view of a flat gradient buffer, or anything else) is about to be written by a kernel that the caller enqueues
		on the compute stream right AFTER this call"""
		reg = self._region(ary._ptr)
		if reg is None:
			return
		# the ranges reported so far have all their writers enqueued: if a bucket is full, send it off now (the range reported
		# by THIS call is not part of it -- its writer is not on the stream yet)
		if reg[5] >= self.bucketBytes:
			self._launch(reg, self._merge(reg[4]))
			reg[3] = self._merge(reg[3] + reg[4])
			reg[4], reg[5] = [], 0
		lo, hi = ary._ptr, ary._ptr + ary.nbytes
		# a reduction of this range still in flight must finish before the range changes under it
		for event, ranges in self.inflight:
			if any(a < hi and lo < b for a, b in ranges):
				check(lib.pz_stream_wait_event(None, event.handle))
		reg[3] = self._subtract(reg[3], (lo, hi))
		reg[4].append([lo, hi])
		reg[5] += hi - lo

	def _launch(self, reg, ranges):
		if not ranges:
			return
		itemsize = reg[2].itemsize
		n = len(ranges)
		ptrs = (ctypes.c_void_p * n)(*[lo for lo, _ in ranges])
		counts = (ctypes.c_int64 * n)(*[(hi - lo) // itemsize for lo, hi in ranges])
		ready = driver.Event(timing=False)
		ready.record()                                   # after the kernels that produced these gradients (compute stream)
		self.stream.waitEvent(ready)
		check(lib.pz_nccl_allreduce_avg_segments(self.comm.handle, dtypeCode(reg[2]), ptrs, counts, n, self.stream.handle))
		done = driver.Event(timing=False)
		done.record(self.stream)
		self.inflight.append((done, ranges))
		self.launches += 1

	def finish(self, tensor):
		"""the reference's sumTensor(name, tensor): on return (in stream order) `tensor` holds the mean over the ranks"""
		key = tensor._ptr
		reg = self.regions.get(key)
		if reg is None or reg[1] != tensor.nbytes:
			# first step with this buffer (or it changed): reduce all of it, remember the region for the next steps
			reg = [key, tensor.nbytes, tensor.dtype, [], [], 0]
			self.regions[key] = reg
			todo = [[key, key + tensor.nbytes]]
		else:
			# everything that is not known to be clean: the complement of the clean ranges
			todo, cursor = [], key
			for lo, hi in self._merge(reg[3]):
				if lo > cursor:
					todo.append([cursor, lo])
				cursor = max(cursor, hi)
			if cursor < key + tensor.nbytes:
				todo.append([cursor, key + tensor.nbytes])
		self._launch(reg, todo)
		for event, _ in self.inflight:                      # join: the update kernel runs after every bucket
			check(lib.pz_stream_wait_event(None, event.handle))
		self.inflight = []
		reg[3], reg[4], reg[5] = [], [], 0                   # next step: nothing is clean until it has been reduced again

	def close(self):
		self.regions, self.inflight = {}, []


class NodeInfo:
	"""reference: Grid.py:38-157 (ParentNode / ChildNode collapse into one symmetric class: NCCL has no parent)"""

	def __init__(self, index, gridsize, device, rendezvous, comm=None):
		self.index = index
		self.gridsize = gridsize
		self.device = device

		self.rendezvous = rendezvous
		self.comm = comm
		self.sync = None

	def attach(self, overlap=True):
		"""Create the NCCL communicator; call after the device of this process has been selected.  `overlap`: average gradient
		buckets on a communication stream while the backward pass is still running (GradientSync)."""
		if self.comm is None and self.gridsize > 1:
			self.comm = NcclCommunicator(self.rendezvous)
		if self.gridsize > 1 and overlap and os.environ.get("PZ_GRID_NO_OVERLAP") is None:
			self.sync = GradientSync(self.comm)
			driver.gradientWriteHook = self.sync.noteWrite
		return self

	def meanValue(self, value):
		# reference: Grid.py:104-111,139-143 -- a python float averaged over the grid
		return self.rendezvous.meanValue(value) if self.gridsize > 1 else value

	def broadcastBuffer(self, name, buffer):
		# reference: Grid.py:114-121,146-150 -- rank 0's bytes replace everybody's
		if self.gridsize > 1:
			driver.flushDeferred()
			self.comm.broadcastBytes(buffer.ptr, buffer.size, root=0)

	def sumTensor(self, name, tensor):
		# reference: Grid.py:123-135,153-157 -- despite the name the result is the MEAN over the grid (beta = 1 / P)
		if self.gridsize > 1:
			if self.sync is not None:
				self.sync.finish(tensor)
			else:
				self.comm.allReduceMean(tensor)

	def sumTensorAndMomentumSGD(self, param, grad, mom, learnRate, momRate):
		if self.gridsize > 1:
			self.comm.allReduceMomentumSGD(param, grad, mom, learnRate, momRate)
		else:
			check(lib.pz_mean_sgd_momentum(dtypeCode(param.dtype), param.ptr, grad.ptr, mom.ptr, param.size, 1.0, learnRate,
										   momRate, None))

	def barrier(self):
		if self.gridsize > 1:
			driver.Device.synchronize()
			self.rendezvous.barrier()

	def close(self):
		if self.sync is not None:
			driver.gradientWriteHook = None
			self.sync.close()
			self.sync = None
		if self.comm is not None:
			self.comm.close()
			self.comm = None
		if self.rendezvous is not None:
			self.rendezvous.close()


def partition(total, gridsize, index):
	"""Contiguous shard [start, stop) of `total` samples for node `index` (reference: TestLib/MultiGPUMnist.py:33-39)."""
	part = total // gridsize
	return index * part, (index + 1) * part


def nodeFromEnvironment():
	"""NodeInfo of a process launched by torchrun / torch.distributed.run (RANK, LOCAL_RANK, WORLD_SIZE, MASTER_*)."""
	size = int(os.environ.get("WORLD_SIZE", "1"))
	rank = int(os.environ.get("RANK", "0"))
	device = int(os.environ.get("LOCAL_RANK", str(rank)))

	if size == 1:
		return NodeInfo(0, 1, device, None)
	return NodeInfo(rank, size, device, TorchRendezvous())


def _nodeRunner(target, index, size, device, port, args, kwargs):
	from . import seam
	seam.install(deviceIdx=device)           # the reference tree over this backend, bound to this node's GPU
	from PuzzleLib import Config
	Config.allowMultiContext = True

	nodeinfo = NodeInfo(index, size, device, SocketRendezvous(index, size, "127.0.0.1", port) if size > 1 else None)
	try:
		driver.Device(device).set()
		nodeinfo.attach()
		target(nodeinfo, *args, **kwargs)
	finally:
		nodeinfo.close()


def runGrid(target, size, *args, devices=None, port=29533, **kwargs):
	"""One process per GPU, like the reference (Grid.py:4-12); `target(nodeinfo, *args, **kwargs)` runs in each."""
	import multiprocessing

	devices = list(range(size)) if devices is None else list(devices)
	ctx = multiprocessing.get_context("spawn")

	nodes = [ctx.Process(target=_nodeRunner, args=(target, index, size, devices[index], port, args, kwargs)) for index in range(size)]
	for node in nodes:
		node.start()
	for node in nodes:
		node.join()

	failed = [index for index, node in enumerate(nodes) if node.exitcode != 0]
	if failed:
		raise RuntimeError("grid nodes %s failed" % failed)
