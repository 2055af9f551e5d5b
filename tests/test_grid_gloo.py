"""CPU tier, world size 2 over gloo: the host side of the data-parallel path -- rendezvous from the torchrun
environment, the 128-byte unique-id broadcast, scalar mean / max / sum, NodeInfo's `sumTensor` (= mean), `broadcastBuffer`
and the fused reduce + momentum update routing, and the batch partition.  The NCCL data plane itself is covered on GPUs
(bench runs at 2 and 8 GPUs)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np


def _free_port():
	with socket.socket() as s:
		s.bind(("127.0.0.1", 0))
		return s.getsockname()[1]


def test_two_ranks_over_gloo():
	worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_grid_worker.py")
	port = _free_port()
	procs = []
	for rank in range(2):
		env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
		procs.append(subprocess.Popen([sys.executable, worker], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
	results = []
	for proc in procs:
		stdout, stderr = proc.communicate(timeout=300)
		assert proc.returncode == 0, stderr[-2000:]
		results.append(json.loads(stdout.strip().splitlines()[-1]))
	results.sort(key=lambda r: r["rank"])

	assert [r["rank"] for r in results] == [0, 1] and [r["device"] for r in results] == [0, 1]
	for r in results:
		assert r["uid_ok"]
		assert r["mean"] == 2.0 and r["max"] == 20.0 and r["sum"] == 3.0
		assert r["params"] == [7.0, 7.0]
		assert r["fused_param"] == [1.25, 1.25]
	want = (np.array(results[0]["grad_local_head"]) + np.array(results[1]["grad_local_head"])) / 2
	for r in results:
		assert np.allclose(r["grad_mean_head"], want, atol=1e-7)
	assert results[0]["shard"] == [0, 65] and results[1]["shard"] == [65, 130]


_SOCKET_WORKER = """
import json, sys
sys.path.insert(0, %r)
from puzzlelib_b200.grid import SocketRendezvous
rank, size, port = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rv = SocketRendezvous(rank, size, "127.0.0.1", port, timeout=120)
out = {"rank": rank}
out["uid"] = rv.broadcastBytes(bytes(range(128)) if rank == 0 else None, root=0) == bytes(range(128))
out["from2"] = rv.broadcastBytes(b"two" if rank == 2 else None, root=2).decode()
out["max"], out["sum"], out["mean"] = rv.maxValue(rank * 1.5), rv.sumValue(rank + 1), rv.meanValue(2.0 * rank)
rv.barrier()
rv.close()
print(json.dumps(out), flush=True)
"""


def test_socket_rendezvous_three_ranks():
	"""the torch-free control plane `runGrid` uses: unique-id broadcast from any root, max / sum / mean, barrier"""
	root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
	port = _free_port()
	procs = [subprocess.Popen([sys.executable, "-c", _SOCKET_WORKER % root, str(rank), "3", str(port)], stdout=subprocess.PIPE,
							  stderr=subprocess.PIPE, text=True) for rank in range(3)]
	results = []
	for proc in procs:
		stdout, stderr = proc.communicate(timeout=300)
		assert proc.returncode == 0, stderr[-2000:]
		results.append(json.loads(stdout.strip().splitlines()[-1]))
	for r in results:
		assert r["uid"] and r["from2"] == "two"
		assert r["max"] == 3.0 and r["sum"] == 6.0 and r["mean"] == 2.0
