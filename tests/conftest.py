import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "liggghts-inl_b200"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _artefacts_built():
    """a fresh checkout has no built artefacts (they are git-ignored): build them once per session"""
    need = [os.path.join(ROOT, "liggghts-inl_b200", "libdem_b200.so"), os.path.join(ROOT, "liggghts-inl_b200", "lmp_b200"),
            os.path.join(ROOT, "oracle", "liboracle.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__ as g
        g.build()
