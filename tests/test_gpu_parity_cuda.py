"""GPU tier: this repository's kernels held DIRECTLY to the outputs of the reference's own CUDA backend.

The seeded case table of tests/golden_cases.py runs through `PuzzleLib.Backend.*` (the reference's function table, the
calls its Modules make) with this backend behind the seam; every output is compared with tests/golden/ref_cuda_ops.npz,
which the SAME table produced on a B200 with PuzzleLib's cuDNN 9 / cuBLAS 12 / NVRTC backend (tools/gen_golden_cuda.py).

Bars (BASELINE.json north_star): bit-exact for integer outputs (max-pool masks, argmax -- ties included) and for pure data
movement; fp32 tensors within 1e-3 relative for the tensor-core contractions (TF32 products here, full fp32 in cuDNN on
this stack), 2e-5 for everything else; 16-bit tensors within 4e-3.
"""
import os

import numpy as np
import pytest

import golden_cases as gc

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CONTRACTIONS = ("conv", "deconv", "gemm")
EXACT = ("memory", "maxpoolmask")


@pytest.fixture(scope="module")
def gold():
	data = np.load(os.path.join(GOLDEN, "ref_cuda_ops.npz"))
	return {key: data[key] for key in data.files}


@pytest.fixture(scope="module")
def table(bnd):
	return gc.bind()


def rel(got, want):
	got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
	return float(np.abs(got - want).max() / (np.abs(want).max() + 1e-30))


@pytest.mark.parametrize("name", sorted(gc.CASES))
def test_case_matches_the_reference_cuda_backend(table, gold, name):
	got = gc.run(table, [name])
	assert got, name
	failures = []

	for key, val in got.items():
		want = gold[key]
		field = key.split("/")[1]
		assert val.shape == want.shape and val.dtype == want.dtype, key

		if field.startswith("in_"):
			assert np.array_equal(val, want), "%s: the seeded inputs differ from the golden file's" % key
		elif val.dtype.kind in "iu" or name.startswith(EXACT):
			if not np.array_equal(val, want):
				failures.append("%s: %d of %d entries differ (must be bit-exact)" % (key, int((val != want).sum()), val.size))
		else:
			if val.dtype == np.float16:
				bar = 4e-3
			elif name.startswith(CONTRACTIONS) and field not in ("bgrad", "bacc", "colsum", "rowsum"):
				bar = 1e-3
			else:
				bar = 2e-5
			err = rel(val, want)
			if not err < bar:
				failures.append("%s: relative error %.3e > %.1e" % (key, err, bar))

	assert failures == [], "\n".join(failures)


@pytest.fixture(scope="module")
def goldSide():
	data = np.load(os.path.join(GOLDEN, "ref_cuda_side.npz"))
	return {key: data[key] for key in data.files}


@pytest.mark.parametrize("name", sorted(gc.SIDE_CASES))
def test_side_module_case_matches_the_reference_cuda_backend(table, goldSide, name):
	"""PReLU, reflection padding, embedding lookup, up-sampling, divisive normalisation against the reference's NVRTC kernels /
	cuDNN (tests/golden/ref_cuda_side.npz): data movement bit-exact, arithmetic within 2e-5 (4e-3 for float16)"""
	got = gc.run(table, [name])
	failures = []
	for key, val in got.items():
		want = goldSide[key]
		field = key.split("/")[1]
		assert val.shape == want.shape and val.dtype == want.dtype, key
		if field.startswith("in_") or val.dtype.kind in "iu" or (field == "y" and ("pad" in name or "nearest" in name or "embed" in name)):
			if not np.array_equal(val, want):
				failures.append("%s: not bit-exact" % key)
		else:
			bar = 4e-3 if val.dtype == np.float16 else 2e-5
			err = rel(val, want)
			if not err < bar:
				failures.append("%s: relative error %.3e > %.1e" % (key, err, bar))
	assert failures == [], "\n".join(failures)


def l2(got, want):
	got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
	return float(np.linalg.norm(got - want) / (np.linalg.norm(want) + 1e-30))


@pytest.mark.parametrize("name,ybar,gbar", [("lenet_n4", 2e-3, 3e-2), ("resnet50_n2", 5e-2, None)])
def test_whole_net_against_the_reference_cuda_backend(table, name, ybar, gbar):
	"""Whole reference models (Models/Nets/LeNet.py, ResNet.py) with seeded parameters, forward + backward, against the same run
	on the reference's CUDA backend.  End-to-end bars are L2 and loose BY NATURE: TF32 rounding (3e-4 per contraction) flips
	max-pool / ReLU decisions on near-ties and train-mode batch norm over 2 images amplifies them over 50 layers; the tight
	per-operator bars are the table above and the teacher-forced per-layer checks of test_gpu_nets.py."""
	data = np.load(os.path.join(GOLDEN, "ref_cuda_nets.npz"))
	got = gc.runNets(table, [name])

	assert list(got[name + "/names"]) == list(data[name + "/names"])
	assert got[name + "/y"].shape == data[name + "/y"].shape
	assert l2(got[name + "/y"], data[name + "/y"]) < ybar
	if gbar is not None:
		assert l2(got[name + "/gradabs"], data[name + "/gradabs"]) < gbar
		for key in data.files:
			if key.startswith(name + "/grad:"):
				assert l2(got[key], data[key]) < 10 * gbar, key
