"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU and
exports every symbol include/cadence_b200.h declares; the Python shims mirror
the reference API and refuse to run on CPU (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_text(name="cadence_b200.h"):
  with open(os.path.join(ROOT, "include", name)) as f:
    return re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)


def _declared_symbols(name="cadence_b200.h"):
  return sorted(set(re.findall(r"\b(cg_[a-z0-9_]+)\s*\(", _header_text(name))))


def _declared_arg_counts(name="cadence_b200.h"):
  """{symbol: number of parameters} parsed from the prototypes of a header."""
  out = {}
  for m in re.finditer(r"\b(cg_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", _header_text(name)):
    args = m.group(2).strip()
    out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
  return out


def test_library_exports_every_declared_symbol():
  from cadence_gemma_b200 import _abi, build
  lib_path = build.build()           # cross-compiles for sm_100a if stale
  lib = ctypes.CDLL(lib_path)
  declared = _declared_symbols()
  assert len(declared) >= 7, declared
  for name in declared:
    assert hasattr(lib, name), f"{name} declared in the header but not exported"
  assert sorted(_abi.SYMBOLS) == declared, "ctypes table out of sync with the header"
  # ... argument for argument (a stale binder would pass the stream in the wrong slot)
  counts = _declared_arg_counts()
  for name, (_, argtypes) in _abi.SYMBOLS.items():
    assert counts[name] == len(argtypes), (name, counts[name], len(argtypes))
  # the experimental header (development taps) is bound separately and is NOT part of the surface
  exp = _declared_symbols("cadence_b200_experimental.h")
  assert sorted(_abi.EXPERIMENTAL_SYMBOLS) == exp
  exp_counts = _declared_arg_counts("cadence_b200_experimental.h")
  for name, (_, argtypes) in _abi.EXPERIMENTAL_SYMBOLS.items():
    assert hasattr(lib, name) and exp_counts[name] == len(argtypes), name
  loaded = _abi.load()
  version = int(re.search(r"#define CG_ABI_VERSION (\d+)", open(os.path.join(ROOT, "include", "cadence_b200.h")).read()).group(1))
  assert loaded.cg_abi_version() == version == _abi.ABI_VERSION
  assert b"workspace" in loaded.cg_status_string(-5)
  assert loaded.cg_scan_workspace_bytes(8, 2048, 2560, 1) > 0
  assert loaded.cg_scan_workspace_bytes(0, 1, 1, 1) == 0
  # fused tensor-core path: host-side shape logic only (no launches without a GPU)
  assert loaded.cg_rglru_fused_supported(2560, 10, 1) == 1      # RecurrentGemma-2B, bf16
  assert loaded.cg_rglru_fused_supported(4096, 16, 1) == 1      # 9B
  assert loaded.cg_rglru_fused_supported(2560, 10, 0) == 0      # fp32 -> scan kernel
  assert loaded.cg_rglru_fused_supported(256, 8, 1) == 0        # head width 32
  # [E/128][2 gates][bw/64][128 x 128 B] + identity [2][128 x 128 B]
  assert loaded.cg_rglru_gate_pack_bytes(2560, 10) == 20 * 2 * 4 * 16384 + 2 * 16384
  assert loaded.cg_rglru_gate_pack_bytes(256, 8) == 0
  assert loaded.cg_rglru_fused_workspace_bytes(8, 2048, 2560) > 3 * 8 * 8 * 64 * 2560
  assert b"fused" in loaded.cg_status_string(-7)


def test_shims_mirror_reference_api_and_have_no_cpu_fallback():
  import cadence_gemma_b200 as cg
  lru = cg.RGLRU(64, 2)
  conv = cg.Conv1D(64, 4)
  assert sorted(lru.state_dict()) == ["a_gate.b", "a_gate.w", "a_param",
                                      "input_gate.b", "input_gate.w"]
  assert sorted(conv.state_dict()) == ["b", "w"]
  assert lru.input_gate.w.shape == (2, 32, 32) and conv.w.shape == (4, 64)
  assert cg.RGLRU.init_cache(3, 64).dtype == torch.float32
  assert cg.Conv1D.init_cache(batch_size=3, width=64, dtype=torch.bfloat16).shape == (3, 3, 64)
  x = torch.randn(1, 8, 64)
  seg = torch.arange(8)[None]
  with torch.no_grad():
    for call in (lambda: lru(x, seg), lambda: conv(x, seg),
                 lambda: cg.rnn_scan(x, x, seg == 0, None)):
      with pytest.raises(RuntimeError, match="no CPU"):
        call()
  # training path: still no CPU fallback; decode steps stay forward-only
  xg = torch.randn(1, 8, 64, requires_grad=True)
  with torch.enable_grad():
    with pytest.raises(RuntimeError, match="no CPU"):
      conv(xg, seg)
    with pytest.raises(RuntimeError, match="no CPU"):
      lru(xg, seg)
    with pytest.raises(RuntimeError, match="forward-only"):
      conv(xg[:, :1], seg[:, :1], torch.zeros(1, 3, 64))


def test_recurrent_block_mirror_state_dict_keys():
  from cadence_gemma_b200.modules import RecurrentBlock
  blk = RecurrentBlock(width=32, num_heads=2, lru_width=64)
  keys = set(blk.state_dict())
  for k in ("linear_x.weight", "linear_y.bias", "linear_out.weight", "conv_1d.w",
            "conv_1d.b", "rg_lru.a_param", "rg_lru.input_gate.w", "rg_lru.a_gate.b"):
    assert k in keys
  cache = RecurrentBlock.init_cache(2, 64, torch.bfloat16)
  assert cache.rg_lru_state.dtype == torch.float32 and cache.conv1d_state.shape == (2, 3, 64)


def test_install_patches_reference_entry_points():
  from oracle import ref_loader
  if not ref_loader.reference_available():
    pytest.skip("reference checkout not mounted")
  from cadence_gemma_b200 import install
  ref = ref_loader.load_reference()
  orig = (ref.layers.rnn_scan, ref.layers.RGLRU.forward, ref.layers.Conv1D.forward)
  install.install(ref.layers, ref.modules)
  try:
    assert ref.layers.rnn_scan is not orig[0]
    assert ref.layers.RGLRU.forward is not orig[1]
    assert ref.layers.Conv1D.forward is not orig[2]
    # state dicts of the reference classes are what the kernels consume
    blk = ref.modules.RecurrentBlock(width=32, num_heads=2, lru_width=64)
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU"):
      blk(torch.randn(1, 4, 32), torch.arange(4)[None])
  finally:
    install.uninstall()
  assert (ref.layers.rnn_scan, ref.layers.RGLRU.forward, ref.layers.Conv1D.forward) == orig


def _integration_md_blocks():
  with open(os.path.join(ROOT, "INTEGRATION.md")) as f:
    text = f.read()
  return re.findall(r"```python\n(.*?)```", text, flags=re.S)


def test_integration_md_stubs_match_the_header():
  """Every `argtypes` list a maintainer would paste from INTEGRATION.md has exactly
  as many entries as the prototype in include/cadence_b200.h (round 1 shipped a
  21-argument stub for a 23-argument function) and names the current ABI version."""
  from cadence_gemma_b200 import _abi
  counts = _declared_arg_counts()
  code = "\n".join(_integration_md_blocks())
  found = re.findall(r"_lib\.(cg_[a-z0-9_]+)\.argtypes\s*=\s*\[(.*?)\]", code, flags=re.S)
  assert len(found) >= 6, [f[0] for f in found]
  for name, body in found:
    n = len([t for t in re.split(r"[,\s]+", body.split("#")[0].strip()) if t])
    assert name in counts, name
    assert n == counts[name] == len(_abi.SYMBOLS[name][1]), (name, n, counts[name])
  assert f"cg_abi_version() == {_abi.ABI_VERSION}" in code
  # every symbol the stubs call exists in the public header (nothing experimental)
  for name in set(re.findall(r"_lib\.(cg_[a-z0-9_]+)", code)):
    assert name in counts, f"INTEGRATION.md uses {name}, which include/cadence_b200.h does not declare"
