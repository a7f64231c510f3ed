import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(autouse=True)
def _jmc_env_switches(monkeypatch):
    """The library caches its JMC_* switches per process (not per launch): re-read them at the start of every test
    (the previous test's monkeypatch has been undone by now) and whenever a test sets one."""
    def reload():
        try:
            import jmcodec_b200
            jmcodec_b200.reload_env()
        except Exception:
            pass
    reload()
    orig = monkeypatch.setenv

    def setenv(name, value, *a, **k):
        orig(name, value, *a, **k)
        if name.startswith("JMC_"):
            reload()
    monkeypatch.setenv = setenv
    yield
