import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden_dir():
    return os.path.join(ROOT, 'tests', 'golden')


def _gpu_ready():
    """(ok, why): the -m gpu tests need a CUDA device of compute capability 10.x.  A missing library on such a box is NOT a
    reason to skip: the product path has no fallback and the tests must fail loudly there."""
    try:
        import torch
        if not torch.cuda.is_available():
            return False, 'no CUDA device'
        if torch.cuda.get_device_capability(0)[0] != 10:
            return False, 'CUDA device is not sm_100 (B200)'
    except Exception as e:          # noqa: BLE001
        return False, f'{type(e).__name__}: {e}'
    return True, ''


def pytest_collection_modifyitems(config, items):
    """Without a B200 the gpu-marked tests are skipped, not failed, so a plain `pytest tests` works everywhere; on the
    GPU box nothing is skipped."""
    gpu_items = [it for it in items if it.get_closest_marker('gpu')]
    if not gpu_items:
        return
    ok, why = _gpu_ready()
    if ok:
        return
    skip = pytest.mark.skip(reason='needs a B200: ' + why)
    for it in gpu_items:
        it.add_marker(skip)
