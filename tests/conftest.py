"""pytest config: registers the `gpu` marker and puts the package directory on sys.path.

The package directory is `megatron-clip_b200/` (not importable by that name because of the hyphen); the
importable package inside it is `clipk`.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "megatron-clip_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu through gpurun)")
