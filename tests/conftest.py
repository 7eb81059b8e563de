import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


CFGS = {"rot_center_depth": 1.0, "fov": 10, "tex_cube_size": 2}  # config.yml:25-27
MIN_DEPTH, MAX_DEPTH = 0.9, 1.1                                      # model.py:49-50


@pytest.fixture(scope="session")
def cfgs():
    return dict(CFGS)
