"""taa_star_b200 — B200-native temporal anti-aliasing resolve (the hot path of cg-tuwien/TAA-STAR).

Layout:
  csrc/       CUDA kernels (sm_100a) and the C-ABI implementation -> libtaa_b200.so
  abi.py      ctypes view of include/taa_b200.h
  host.py     TaaContext (raw resolve / frame calls) and Taa (mirror of `class taa<CF>`, source/taa.hpp)
  synth.py    synthetic jittered G-buffer sequences (the renderer is out of scope)
  sharded.py  row-band sharding across GPUs with history-halo exchange
"""
from . import abi  # noqa: F401

__all__ = ["abi"]
