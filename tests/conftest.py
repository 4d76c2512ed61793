"""pytest configuration: marker registration and shared fixtures.

`-m "not gpu"`  : oracle vs golden vectors, host logic, C-ABI symbol table (CPU only).
`-m gpu`        : parity of the CUDA path (through the C-ABI) against the oracle.
"""
import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "golden_ref_strict.npz"))
    return {k: z[k] for k in z.files}


def crc(*arrays):
    import zlib
    c = 0
    for a in arrays:
        c = zlib.crc32(np.ascontiguousarray(a).tobytes(), c)
    return np.uint32(c)


def unpack_mask(bits, shape):
    n = int(np.prod(shape))
    return np.unpackbits(bits)[:n].reshape(shape).astype(np.int32)
