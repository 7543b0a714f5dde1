"""Places the UNMODIFIED reference package under baseline/_ref (git-ignored; it
travels to the GPU box with the repo snapshot).

    python baseline/install_reference.py [--reference /root/reference]

First tries the offline pip install the task contract names:

    python -m pip install --no-index --no-build-isolation --no-deps \
        --find-links /opt/wheelhouse --target baseline/_ref <copy of the reference>

The reference's build backend is poetry-core (pyproject.toml:40-42), which is not
in the offline wheelhouse, so that install fails here ("No module named
'poetry'").  The package is pure Python (no native sources),
so what pip would have placed under --target is exactly the `recurrentgemma/`
package directory: this script then copies that directory, byte for byte,
instead.  Nothing under baseline/_ref is ever edited, and nothing of it is
tracked by git.

Used by: `bench.py --impl reference` (the reference arm runs the reference's own
Conv1D.forward / RGLRU.forward on the host cores) and the `-m gpu` install tests
(the reference's own RecurrentBlock / ResidualBlock / Griffin classes with our
kernels patched in).  The product path never imports it.
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")


def install(reference: str = "/root/reference", verbose: bool = False) -> str | None:
  """Returns "pip" / "copy", or None if the reference checkout is absent."""
  if not os.path.isfile(os.path.join(reference, "recurrentgemma", "torch", "layers.py")):
    return None
  marker = os.path.join(TARGET, "recurrentgemma", "torch", "layers.py")
  if os.path.isfile(marker):
    return "present"
  os.makedirs(TARGET, exist_ok=True)
  how = None
  with tempfile.TemporaryDirectory() as tmp:
    src = os.path.join(tmp, "reference")
    shutil.copytree(reference, src, ignore=shutil.ignore_patterns(".git", "*.ipynb"))
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
           "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode == 0 and os.path.isfile(marker):
      how = "pip"
    else:
      if verbose:
        sys.stderr.write(proc.stdout[-2000:] + proc.stderr[-2000:])
      dst = os.path.join(TARGET, "recurrentgemma")
      if os.path.isdir(dst):
        shutil.rmtree(dst)
      shutil.copytree(os.path.join(reference, "recurrentgemma"), dst,
                      ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
      how = "copy"
  with open(os.path.join(TARGET, "INSTALLED_BY"), "w") as f:
    f.write(f"baseline/install_reference.py: {how} from {reference}\n")
  return how


if __name__ == "__main__":
  ap = argparse.ArgumentParser()
  ap.add_argument("--reference", default="/root/reference")
  a = ap.parse_args()
  print(install(a.reference, verbose=True))
