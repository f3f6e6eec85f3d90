"""In-tree build of the CUDA product library (csrc/libslb.so, sm_100a) with the committed Makefile.
nvcc cross-compiles without a GPU, so this runs in the CPU-only build container too."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(CSRC, "libslb.so")


def build(force=False, verbose=False):
    cmd = ["make", "-C", CSRC, "-j8"] + (["-B"] if force else [])
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    return LIB
