"""One rank of the world-size-2 gloo test (tests/test_grid_gloo.py): exercises the host side of puzzlelib_b200.grid -- the
rendezvous, NodeInfo's routing of the reference's `sumTensor / broadcastBuffer / meanValue` surface, the shard partition --
with a gloo-backed stand-in for the NCCL communicator (no GPU in the CPU tier).  Prints one JSON line."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


class GlooCommunicator:
	"""Same methods as grid.NcclCommunicator, over torch.distributed (gloo) on host arrays."""

	def __init__(self, rendezvous):
		import torch
		self.torch, self.dist = torch, rendezvous.dist
		self.rank, self.size = rendezvous.rank, rendezvous.size

	def allReduceMean(self, ary):
		t = self.torch.from_numpy(ary.host)
		self.dist.all_reduce(t)
		t /= self.size

	def broadcastBytes(self, ptr, nbytes, root=0):
		t = self.torch.from_numpy(ptr.view(np.uint8).reshape(-1)[:nbytes])
		self.dist.broadcast(t, src=root)

	def allReduceMomentumSGD(self, param, grad, mom, learnRate, momRate):
		self.allReduceMean(grad)
		mom.host[...] = momRate * mom.host + learnRate * grad.host
		param.host[...] += mom.host

	def close(self):
		pass


class HostArray:
	"""What NodeInfo touches of a GPUArray: .ptr, .size, .dtype (here the "pointer" is the numpy array itself)."""

	def __init__(self, host):
		self.host = host
		self.ptr, self.size, self.dtype = host, host.nbytes, host.dtype


def main():
	from puzzlelib_b200 import grid

	node = grid.nodeFromEnvironment()                       # RANK / WORLD_SIZE / MASTER_* as torchrun sets them
	assert node.gridsize == 2 and node.index == int(os.environ["RANK"])
	node.comm = GlooCommunicator(node.rendezvous)
	out = {"rank": node.index, "device": node.device}

	payload = node.rendezvous.broadcastBytes(bytes(range(128)) if node.index == 0 else None, root=0)
	out["uid_ok"] = payload == bytes(range(128))             # the 128-byte ncclUniqueId travels this way

	out["mean"] = node.meanValue(1.0 + 2.0 * node.index)    # 1 and 3 -> 2
	out["max"] = node.rendezvous.maxValue(10.0 * (node.index + 1))
	out["sum"] = node.rendezvous.sumValue(node.index + 1)

	rng = np.random.RandomState(100 + node.index)
	gradient = rng.randn(1000).astype(np.float32)
	out["grad_local_head"] = gradient[:4].tolist()
	node.sumTensor("grad", HostArray(gradient))              # despite the name: the MEAN over the grid (Grid.py:123-135)
	out["grad_mean_head"] = gradient[:4].tolist()

	params = np.full(16, float(node.index + 7), dtype=np.float32)
	node.broadcastBuffer("params", HostArray(params))        # rank 0's bytes replace everybody's
	out["params"] = params[:2].tolist()

	p, g, m = np.ones(8, np.float32), np.full(8, float(node.index), np.float32), np.zeros(8, np.float32)
	node.sumTensorAndMomentumSGD(HostArray(p), HostArray(g), HostArray(m), 0.5, 0.9)
	out["fused_param"] = p[:2].tolist()                      # 1 + 0.5 * mean(0, 1) = 1.25

	out["shard"] = list(grid.partition(130, node.gridsize, node.index))
	node.rendezvous.barrier()
	node.comm = None                                         # nothing NCCL to destroy
	node.close()
	print(json.dumps(out), flush=True)


if __name__ == "__main__":
	main()
