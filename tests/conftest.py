import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))          # oracle modules: TEST-ONLY imports
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    # a fresh checkout has no built library (it is git-ignored): build it once, like __graft_entry__.build() does
    lib = os.path.join(ROOT, "zk_symmetric_crypto_b200", "libs2c_b200.so")
    if not os.path.exists(lib):
        import shutil
        import subprocess
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            env = dict(os.environ, PATH=os.environ.get("PATH", "") + ":/usr/local/cuda/bin")
            subprocess.run(["make", "-j8", "-C", ROOT], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def backend():
    import zk_symmetric_crypto_b200 as z
    be = z.Backend(0)
    yield be
    be.close()
