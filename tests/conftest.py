import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(autouse=True)
def _in_tmp_cwd(tmp_path, monkeypatch):
    """The estimators create results/<algo>/<main_dir>/... relative to the cwd (like the reference)."""
    monkeypatch.chdir(tmp_path)


@pytest.fixture(scope='session', autouse=True)
def _built_library():
    """libdae_sm100.so is git-ignored: build it (nvcc cross-compiles sm_100a without a GPU; cached objects make this a no-op)."""
    from dae_rnn_news_recommendation_b200 import build, _cabi
    if not _cabi.LIB_PATH.exists():
        build.build()
