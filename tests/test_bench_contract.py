"""The driver-facing contract of bench.py that can be checked without a GPU:
the reference arm (`--impl reference`, the CPU restatement of the reference's
path, the one place outside tests/ that may execute oracle/) prints ONE JSON
line with the agreed keys, non-zero ranks of a torchrun launch stay silent, and
the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=300):
  e = dict(os.environ)
  e.update(env or {})
  return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=e,
                        capture_output=True, text=True, timeout=timeout)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
  r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
  assert r.returncode == 0, r.stderr[-2000:]
  lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
  assert len(lines) == 1, lines
  d = json.loads(lines[0])
  assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1
  assert d["metric"] == "rglru_conv1d_prefill_tokens_per_sec" and d["unit"] == "tokens/s"
  assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
  assert "workload" in d["config"] and "model" not in d["config"]
  assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
  assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
  assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                      "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_silently():
  r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
           env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
  assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_arm_fails_loudly_without_a_gpu():
  r = _run(["--steps", "1", "--warmup", "1"])
  assert r.returncode != 0
  assert "{" not in r.stdout, "no bench line may be printed without the CUDA path"
