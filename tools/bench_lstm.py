"""BASELINE.json configs[4]: Modules/RNN.py LSTM hidden=1024 seq=256 batch=64 fp32 fwd+bwd on one B200.
Times the eager module API and a CUDA-graph replay of the same step; prints ms/step and achieved TFLOP/s
(824.6 GFLOP per fwd+bwd step).  usage: python tools/bench_lstm.py [T] [B] [H]"""
import json
import sys
import numpy as np
sys.path.insert(0, ".")
from puzzlelib_b200 import driver
from puzzlelib_b200 import seam

seam.install()                      # the reference's Modules/RNN.py over this backend
import PuzzleLib.Modules as M                      # noqa: E402
from PuzzleLib.Backend import gpuarray             # noqa: E402


def backend():
	return gpuarray.backend


def main():
	T, B, H = (int(v) for v in (sys.argv[1:4] + ["256", "64", "1024"][len(sys.argv) - 1:]))
	backend()
	np.random.seed(1)
	rnn = M.RNN(H, H, layers=1, mode="lstm", getSequences=True, initscheme="xavier")
	x = gpuarray.to_gpu(np.random.randn(T, B, H).astype(np.float32))
	g = gpuarray.to_gpu(np.random.randn(T, B, H).astype(np.float32))

	def step():
		rnn.zeroGradParams()
		rnn(x)
		rnn.backward(g)
		rnn.reset()

	def timeit(fn, sync, reps=5):
		fn()
		sync()
		e0, e1 = driver.Event(), driver.Event()
		e0.record(None)
		for _ in range(reps):
			fn()
		e1.record(None)
		e1.synchronize()
		return e0.timeTill(e1) / reps

	for _ in range(2):
		step()
	eager = timeit(step, driver.Device.synchronize)
	flops = 3 * 2.0 * T * B * (H * 4 * H + H * 4 * H)
	out = {"config": "LSTM T=%d B=%d I=H=%d fp32 fwd+bwd" % (T, B, H), "eager_ms": eager, "eager_tflops": flops / eager / 1e9}
	try:
		graph = driver.StepGraph(step, warmup=1)
		e0, e1 = driver.Event(), driver.Event()
		graph.launch()
		graph.synchronize()
		e0.record(graph.stream)
		for _ in range(5):
			graph.launch()
		e1.record(graph.stream)
		e1.synchronize()
		ms = e0.timeTill(e1) / 5
		out.update({"graph_ms": ms, "graph_tflops": flops / ms / 1e9})
	except Exception as e:      # noqa: BLE001
		out["graph"] = "unavailable: %s" % e
	print(json.dumps(out))


if __name__ == "__main__":
	main()
