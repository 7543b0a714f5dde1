"""Tiny (de)serialiser for golden fixtures: dict[str, torch.Tensor] <-> .npz.

bf16 has no numpy dtype, so bf16 tensors are stored as their raw uint16 words
under ``<key>::bf16``; bool as uint8 under ``<key>::bool``.
"""
from __future__ import annotations

import os

import numpy as np
import torch

GOLDEN_DIR = os.path.dirname(os.path.abspath(__file__))


def save(name: str, tensors: dict) -> str:
  arrays = {}
  for key, value in tensors.items():
    if value is None:
      continue
    t = value.detach().cpu().contiguous()
    if t.dtype == torch.bfloat16:
      arrays[key + "::bf16"] = t.view(torch.int16).numpy().view(np.uint16)
    elif t.dtype == torch.bool:
      arrays[key + "::bool"] = t.to(torch.uint8).numpy()
    else:
      arrays[key] = t.numpy()
  path = os.path.join(GOLDEN_DIR, name + ".npz")
  np.savez_compressed(path, **arrays)
  return path


def load(name: str) -> dict:
  path = os.path.join(GOLDEN_DIR, name + ".npz")
  out = {}
  with np.load(path) as data:
    for key in data.files:
      arr = data[key]
      if key.endswith("::bf16"):
        out[key[:-6]] = torch.from_numpy(
            arr.view(np.int16).copy()).view(torch.bfloat16)
      elif key.endswith("::bool"):
        out[key[:-6]] = torch.from_numpy(arr.copy()).to(torch.bool)
      else:
        out[key] = torch.from_numpy(arr.copy())
  return out


def cases(prefix: str) -> list[str]:
  return sorted(
      f[:-4] for f in os.listdir(GOLDEN_DIR)
      if f.startswith(prefix) and f.endswith(".npz"))
