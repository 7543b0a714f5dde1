"""pytest configuration: registers the `gpu` marker and puts the repo on sys.path."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line(
      "markers", "gpu: needs a CUDA device (run with `-m gpu` on the B200 box)")


def pytest_collection_modifyitems(config, items):
  import torch
  if torch.cuda.is_available():
    return
  skip = pytest.mark.skip(reason="no CUDA device")
  for item in items:
    if "gpu" in item.keywords:
      item.add_marker(skip)


@pytest.fixture(autouse=True)
def _inference_by_default():
  """Parity tests exercise the inference kernels: autograd is off unless a test
  turns it on (``torch.enable_grad()``) -- with grad enabled the modules take
  their differentiable training path instead."""
  import torch
  with torch.no_grad():
    yield
