"""Domain decomposition LOGIC on the CPU: the brick cases of tests/brick_cases.py on the emulator build
(kernels compiled by g++ against tests/emu/cuda_emu.h, ranks talk through an in-process mailbox).  The
NCCL build of the same code runs the same cases in tests/test_gpu_z_bricks.py on the GPU box."""
import pytest

import azt
import brick_cases


@pytest.fixture(scope="module")
def emu():
    return azt.emu_lib()


@pytest.mark.parametrize("name,case", brick_cases.ALL_CASES, ids=[c[0] for c in brick_cases.ALL_CASES])
def test_emulated_bricks(emu, name, case):
    case(emu)
