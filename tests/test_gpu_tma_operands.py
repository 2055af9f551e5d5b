"""The copy-engine operand paths of the conv engine (MODE_K_POS_TMA, MODE_MN_TMA; csrc/pz_conv.cu: PZ_TMA_WGRAD, PZ_TMA_FPROP) against a
float64 contraction of the same inputs, next to the rounding producers they replace.  The levels are read once per process, so every
setting runs tools/check_tma_*.py in its own interpreter; the scripts print the worst error relative to the largest element and
"OK" when it is inside the bound `north_star` states for fp32 tensors (1e-3; 2e-3 for half storage)."""
import os, subprocess, sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(script, env, *args):
	e = dict(os.environ)
	e.update(env)
	r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", script), *args], cwd=ROOT, env=e, capture_output=True, text=True, timeout=300)
	assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
	last = r.stdout.strip().splitlines()[-1]
	assert last.startswith("worst") and last.endswith("OK"), r.stdout[-2000:]
	return float(last.split()[1])


@pytest.mark.gpu
@pytest.mark.parametrize("level", ["0", "1", "2"])
def test_wgrad_operands_through_the_copy_engine_float32(level):
	worst = run("check_tma_wgrad.py", {"PZ_TMA_WGRAD": level}, "f32")
	# copied tiles are rounded in place to the value the gathering producers round to: every level stays near 3e-4
	# (un-rounded, i.e. truncated by the tensor core, two copied operands reached 8.5e-4)
	assert worst < 5e-4


@pytest.mark.gpu
@pytest.mark.parametrize("level", ["0", "2"])
def test_wgrad_operands_through_the_copy_engine_float16(level):
	run("check_tma_wgrad.py", {"PZ_TMA_WGRAD": level}, "f16")


@pytest.mark.gpu
@pytest.mark.parametrize("level", ["0", "1", "2"])
def test_mn_major_activation_operand_through_the_copy_engine(level):
	worst = run("check_tma_fprop.py", {"PZ_TMA_FPROP": level})
	assert worst < 5e-4
