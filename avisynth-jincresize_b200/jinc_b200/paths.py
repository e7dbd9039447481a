"""Locations of the in-tree native artefacts (all built by the top-level Makefile)."""
import os

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPO = os.path.dirname(PKG_DIR)


def minihost_lib() -> str:
    return os.path.join(REPO, "minihost", "libavs_minihost.so")


def oracle_lib() -> str:
    return os.path.join(REPO, "oracle", "libjinc_oracle.so")


def ref_plugin() -> str:
    """The unmodified reference compiled from /root/reference (oracle/_ref; prebuilt copy travels to the GPU box)."""
    return os.path.join(REPO, "oracle", "_ref", "libjincresize_ref.so")


def cuda_lib() -> str:
    return os.path.join(PKG_DIR, "libjinc_b200.so")


def b200_plugin() -> str:
    return os.path.join(PKG_DIR, "libjincresize_b200.so")


def fma_peak_tool() -> str:
    return os.path.join(PKG_DIR, "fma_peak")
