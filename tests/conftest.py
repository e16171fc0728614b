import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle_binding import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ab():
    import accessor_blas_b200
    return accessor_blas_b200


@pytest.fixture(scope="session")
def handle(ab):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return ab.Handle(0)


@pytest.fixture(scope="session")
def refk():
    """The reference's own CUDA kernels, if oracle/_ref was built and shipped."""
    from oracle_binding import REF_LIB, RefKernels
    if not REF_LIB.exists():
        pytest.skip("oracle/_ref/libref_kernels.so not present")
    return RefKernels()
