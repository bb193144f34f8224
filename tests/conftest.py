import os
import sys
import warnings

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
	sys.path.insert(0, ROOT)


def pytest_configure(config):
	config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(autouse=True)
def _quiet():
	with warnings.catch_warnings():
		warnings.simplefilter("ignore")
		yield


@pytest.fixture
def twin():
	"""Route the product's Python classes to the host twin of the engine (CPU tests only)."""
	from tests.twin import use_twin
	with use_twin() as state:
		yield state
