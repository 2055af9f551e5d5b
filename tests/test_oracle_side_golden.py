"""CPU tier: the oracle's side-module functions (PReLU, reflection padding, embedding lookup, up-sampling, divisive normalisation)
held to the outputs of the reference's own CUDA backend -- tests/golden/ref_cuda_side.npz, written on a B200 by
`tools/gen_golden_cuda.py --impl ref --side` from the seeded table tests/golden_cases.py SIDE_CASES."""
import os

import numpy as np
import pytest

from oracle import ops
import golden_cases as gc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_cuda_side.npz")


@pytest.fixture(scope="module")
def gold():
	data = np.load(GOLDEN)
	return {key: data[key] for key in data.files}


def case(gold, name):
	prefix = name + "/"
	out = {key[len(prefix):]: val for key, val in gold.items() if key.startswith(prefix)}
	assert out, "golden file has no case %s" % name
	return out


def close(got, want, tol=2e-5):
	got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
	return float(np.abs(got - want).max()) <= tol * (np.abs(want).max() + 1e-30)


def test_golden_file_covers_the_table(gold):
	assert {key.split("/")[0] for key in gold} == set(gc.SIDE_CASES)


@pytest.mark.parametrize("name,shared", [("side_prelu", False), ("side_prelu_shared", True)])
def test_prelu(gold, name, shared):
	c = case(gold, name)
	assert close(ops.prelu(c["in_x"], c["in_slopes"], shared), c["y"])
	dx, ds = ops.prelu_bwd(c["in_x"], c["in_dy"], c["in_slopes"], shared)
	assert close(dx, c["dx"]) and close(ds, c["dslopes"].ravel())


@pytest.mark.parametrize("name,pad", [("side_pad1d_f32", (2, 3)), ("side_pad2d_f32", (2, 1, 3, 0)), ("side_pad2d_f16", (1, 2, 2, 2))])
def test_reflection_padding(gold, name, pad):
	c = case(gold, name)
	assert np.array_equal(ops.reflectpad(c["in_x"], pad), c["y"])                      # pure data movement
	assert close(ops.reflectpad_bwd(c["in_dy"], pad), c["dx"], 4e-3 if name.endswith("f16") else 2e-5)


def test_embedding(gold):
	c = case(gold, "side_embed_f32")
	assert np.array_equal(ops.embed(c["in_idx"], c["in_W"]), c["y"])
	assert close(ops.embed_bwd(c["in_idx"], c["in_dy"], c["in_W"], 0.25), c["W_after"])


@pytest.mark.parametrize("name,scale", [("side_up2d_nearest", (2, 3)), ("side_up3d_nearest", (2, 1, 2))])
def test_upsample_nearest(gold, name, scale):
	c = case(gold, name)
	assert np.array_equal(ops.upsample_nearest(c["in_x"], scale), c["y"])
	assert close(ops.upsample_nearest_bwd(c["in_dy"], scale), c["dx"])


@pytest.mark.parametrize("name,scale", [("side_up2d_linear", (2, 3)), ("side_up3d_linear", (2, 2, 3)), ("side_up3d_linear_hw", (2, 2, 2))])
def test_upsample_linear(gold, name, scale):
	c = case(gold, name)
	assert close(ops.upsample_linear(c["in_x"], scale), c["y"])
	assert close(ops.upsample_linear_bwd(c["in_dy"], scale), c["dx"])


def test_the_3d_linear_addressing_quirk_is_real(gold):
	"""with inh != inw the reference's 3-d forward kernel reads one tap from another row (Upsample.py:241): the golden output matches
	the oracle WITH the quirk and differs from the mathematically clean interpolation"""
	c = case(gold, "side_up3d_linear_hw")
	assert not close(ops.upsample_linear(c["in_x"], (2, 2, 2), quirk=False), c["y"], 1e-3)


def test_divisive_normalisation(gold):
	c = case(gold, "side_lcn")
	x, dy, means = c["in_x"], c["in_dy"], c["means"]
	assert close(ops.pool2d(x, 5, 1, 2, "avgWithPad"), means)
	assert close(ops.lcn(x, means, 5, 1e-2, 0.75, 2.0), c["y"])
	dx, dmeans = ops.lcn(x, means, 5, 1e-2, 0.75, 2.0, grad=dy)
	assert close(dx, c["dx"]) and close(dmeans, c["dmeans"])
