"""Locations of the product's in-tree native artefacts (built by the top-level Makefile)."""
import os

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPO = os.path.dirname(PKG_DIR)


def cuda_lib() -> str:
    """sm_100a kernels + the C ABI of include/jinc_b200.h"""
    return os.path.join(PKG_DIR, "libjinc_b200.so")


def b200_plugin() -> str:
    """the AviSynth+ C plugin"""
    return os.path.join(PKG_DIR, "libjincresize_b200.so")


def vs_plugin() -> str:
    """the VapourSynth (API 4) plugin"""
    return os.path.join(PKG_DIR, "libvsjincresize_b200.so")


def fma_peak_tool() -> str:
    return os.path.join(PKG_DIR, "fma_peak")
