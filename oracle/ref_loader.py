"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the *unmodified* reference hot-path modules from ``/root/reference``
(present in the build container only) or, where that does not exist -- the GPU
box --, from the git-ignored copy ``baseline/_ref`` that
``baseline/install_reference.py`` makes in the build container and that travels
with the repo snapshot, so that

  * ``tests/golden/make_golden.py`` can generate golden input/output vectors,
  * CPU tests can cross-check the oracle restatements against the real code,
  * the ``-m gpu`` install tests can drive the reference's OWN ``RecurrentBlock`` /
    ``ResidualBlock`` / ``Griffin`` classes with the kernels patched in,
  * ``bench.py --impl reference`` can time the reference's own CPU path.

``import recurrentgemma`` itself fails here because
``recurrentgemma/__init__.py:18-20`` pulls in ``timm`` through
``vit/dino_siglip.py:3``; the parent packages are therefore registered as empty
stubs whose ``__path__`` points at the reference tree, which lets
``recurrentgemma.torch.layers`` / ``.modules`` / ``recurrentgemma.common`` load
without executing either ``__init__.py``.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root() -> str:
  env = os.environ.get("CADENCE_REFERENCE_ROOT")
  if env:
    return env
  for cand in ("/root/reference", os.path.join(_REPO, "baseline", "_ref")):
    if os.path.isfile(os.path.join(cand, "recurrentgemma", "torch", "layers.py")):
      return cand
  return "/root/reference"


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
  return os.path.isfile(
      os.path.join(REFERENCE_ROOT, "recurrentgemma", "torch", "layers.py"))


def load_reference():
  """Returns a namespace with ``common``, ``layers`` and ``modules``."""
  if not reference_available():
    raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
  import torch
  from torch import nn

  pkg_root = os.path.join(REFERENCE_ROOT, "recurrentgemma")
  for name, path in (
      ("recurrentgemma", pkg_root),
      ("recurrentgemma.torch", os.path.join(pkg_root, "torch")),
  ):
    if name not in sys.modules:
      stub = types.ModuleType(name)
      stub.__path__ = [path]
      sys.modules[name] = stub
  if "recurrentgemma.vit" not in sys.modules:
    # griffin.py imports the timm-backed encoder; give it an inert stand-in.
    vit = types.ModuleType("recurrentgemma.vit")

    class VisionEncoder(nn.Module):  # pylint: disable=missing-class-docstring

      def __init__(self, *args, **kwargs):
        super().__init__()

      def forward(self, *args, **kwargs):
        raise RuntimeError("vision encoder is stubbed out in the oracle loader")

    vit.VisionEncoder = VisionEncoder
    sys.modules["recurrentgemma.vit"] = vit

  ns = types.SimpleNamespace()
  ns.common = importlib.import_module("recurrentgemma.common")
  ns.layers = importlib.import_module("recurrentgemma.torch.layers")
  ns.modules = importlib.import_module("recurrentgemma.torch.modules")
  ns.torch = torch
  return ns


def load_reference_griffin():
  """Additionally imports ``recurrentgemma.torch.griffin`` (B=1 only, D4)."""
  ns = load_reference()
  ns.griffin = importlib.import_module("recurrentgemma.torch.griffin")
  return ns
