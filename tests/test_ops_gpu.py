"""GPU parity tests for every kernel behind the C ABI (CUDA path vs the CPU oracle)."""
import pytest

pytestmark = pytest.mark.gpu


def _names():
    # the table itself imports torch.cuda-free modules only; safe to import on CPU for collection
    from tests.gpu_checks import CHECKS
    return sorted(CHECKS)


@pytest.mark.parametrize("name", _names())
def test_kernel_parity(name):
    from tests.gpu_checks import CHECKS
    res = CHECKS[name]()
    assert res, name
