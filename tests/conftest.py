import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
if str(ROOT / "tests") not in sys.path:
    sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_util

    return oracle_util.Oracle()


@pytest.fixture(scope="session")
def ctx():
    import pb_starphase_b200 as sp

    c = sp.Context(int(os.environ.get("LOCAL_RANK", "0")))
    yield c
    c.close()
