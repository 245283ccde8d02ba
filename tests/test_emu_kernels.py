"""Kernel LOGIC tests without a GPU: the same .cu/.cuh sources compiled by g++ against the
host-fiber stand-in of tests/emu (see cuda_emu.h) and checked against the oracle.  These do not
replace the `-m gpu` parity tests (tests/test_gpu_parity.py run the nvcc build on a B200); they
catch indexing / synchronisation mistakes before a GPU round trip."""
import pytest

import azt
import parity_cases


@pytest.fixture(scope="module")
def emu():
    return azt.emu_lib()


@pytest.mark.parametrize("name,case", parity_cases.ALL_CASES, ids=[c[0] for c in parity_cases.ALL_CASES])
def test_emulated_kernels_match_oracle(emu, name, case):
    case(emu)
