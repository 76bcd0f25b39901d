"""raytracert_b200 -- B200-native render hot path of the TU Delft TI1805 ray tracer.

The product is native: ``csrc/`` (hand-written sm_100a CUDA behind the C ABI of ``include/rt_b200.h``)
and ``host/`` (C++ host side in the reference's own style).  This Python package is only the thin
ctypes glue the tests and ``bench.py`` use to call that native code; it contains no compute and no
fallback -- importing :mod:`raytracert_b200.binding` fails loudly if ``librt_b200.so`` is missing.
"""
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
BUILD_DIR = os.path.join(PKG_DIR, "_build")
