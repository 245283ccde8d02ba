"""Domain decomposition on real GPUs: the brick cases of tests/brick_cases.py with NCCL between the
devices of one box.  Needs >= 2 GPUs (gpurun --gpus 2/4/8); cases that need more devices than the box
has are skipped -- loudly, with the count in the reason.  (The file name sorts behind test_gpu_parity.py on purpose:
the single-GPU parity tests run first.)"""
import pytest

import brick_cases
from aztotmd_b200 import abi

pytestmark = pytest.mark.gpu

NEED = {"eight_bricks_2x2x2_ionic": 8, "four_bricks_2x2": 4, "distributed_input_with_ids": 4,
        "bricks_rdf_histogram_is_global_and_exact": 4}


@pytest.fixture(scope="module")
def lib():
    return abi.load_library()


@pytest.mark.parametrize("name,case", brick_cases.ALL_CASES, ids=[c[0] for c in brick_cases.ALL_CASES])
def test_bricks_over_nccl(lib, name, case):
    need = NEED.get(name, 2)
    have = lib.az_device_count()
    if name == "brick_layout_errors":
        need = 1
    if have < need:
        pytest.skip("needs %d GPUs, this box has %d" % (need, have))
    # run_bricks skips (instead of waiting for ranks that can never join the communicator) if a case asks for more ranks
    # than devices are listed here
    case(lib, devices=list(range(have)))
