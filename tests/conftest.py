import json, os, sys
import pytest
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path: sys.path.insert(0, ROOT)

def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")

def hx(s): return bytes.fromhex(s[2:] if s.startswith("0x") else s)

@pytest.fixture(scope="session")
def eth():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "eth_vectors.json")))

@pytest.fixture(scope="session")
def pyv():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "pyref_vectors.json")))
