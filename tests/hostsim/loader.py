"""Builds (g++) and loads the TEST-ONLY host harness around the kernels' math headers."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_hostsim.so")
SRC = os.path.join(HERE, "hostsim.cpp")
CSRC = os.path.join(HERE, "..", "..", "casualhdrsplat_b200", "csrc")


def load():
    deps = [SRC, os.path.join(CSRC, "chs_math.cuh"), os.path.join(CSRC, "chs_spline.cuh"), os.path.join(CSRC, "chs_sh.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", SO, SRC])
    return ctypes.CDLL(SO)
