"""Data-parallel gradient synchronisation: the `NodeInfo` of the reference's Grid.py re-done over NCCL.

The reference runs one process per GPU and averages the flat gradient buffer through a parent/child STAR over CUDA-IPC
mapped memory with `multiprocessing.SimpleQueue`s as the control plane (reference: Grid.py:4-157).  Here the process
model and the NodeInfo surface stay -- `sumTensor(name, tensor)`, `broadcastBuffer(name, buffer)`, `meanValue(value)`,
`index`, `gridsize`, `device`, `close()` -- so `Optimizer` is untouched, while the data plane is ONE ncclAllReduce /
ncclBroadcast per call over NVLink / NVSwitch and the control plane (exchange of the 128-byte ncclUniqueId, scalar
means, barriers) is a `Rendezvous` object.  `TorchRendezvous` rides on torch.distributed's gloo backend (present in the
image and what `torchrun` sets up for bench.py); torch is plumbing here -- no tensor of the hot path ever enters it.
"""
import ctypes
import os
from ctypes import byref

from . import driver
from .driver import lib, check, dtypeCode


class TorchRendezvous:
	"""Host-side control plane over torch.distributed (gloo): bytes broadcast, barrier, float max / mean."""

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


class NodeInfo:
	"""reference: Grid.py:38-157 (ParentNode / ChildNode collapse into one symmetric class: NCCL has no parent)"""

	def __init__(self, index, gridsize, device, rendezvous, comm=None):
		self.index = index
		self.gridsize = gridsize
		self.device = device

		self.rendezvous = rendezvous
		self.comm = comm

	def attach(self):
		"""Create the NCCL communicator; call after the device of this process has been selected."""
		if self.comm is None and self.gridsize > 1:
			self.comm = NcclCommunicator(self.rendezvous)
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

	nodeinfo = NodeInfo(index, size, device, TorchRendezvous(index, size, "127.0.0.1", port) if size > 1 else None)
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
