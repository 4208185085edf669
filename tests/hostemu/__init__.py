"""CPU emulation of the device arithmetic (TEST INFRASTRUCTURE ONLY).

Builds ``tests/hostemu/emu.cpp`` — which includes the TBK_HD headers of
``pythtb_b200/csrc`` — with g++ and loads it through ctypes.  The product
package never imports this.
"""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "pythtb_b200", "csrc")
SO = os.path.join(HERE, "libtbk_hostemu.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "emu.cpp")
    deps = [src] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return SO
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-ffp-contract=off", "-I", CSRC, "-x", "c++", src, "-o", SO]
    subprocess.check_call(cmd)
    return SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib
