"""GPU: the whole product stack behind the reference's call - `Device / Robot / OSC.generate(targets)` on a fake
mujoco_py-shaped simulator, `irlosc_step_host` underneath - must return the forces the unmodified reference returned
for every instance of the golden files (or raise its IndexError).  Same body as tests/test_dropin_golden.py, which
runs it on the CPU with the host build of the kernel; written after round 1's GPU budget was spent, collected last.
"""
import pytest

from conftest import GOLDEN_CASES, GOLDEN_CASES_F4
from test_dropin_golden import generate_on_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", GOLDEN_CASES + GOLDEN_CASES_F4)
def test_generate_on_the_gpu_reproduces_the_reference_goldens(case):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    generate_on_golden(case)
