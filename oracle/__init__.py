"""CPU oracle for the IC-GVINS hot paths.  TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import or execute anything in this directory."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libicg_oracle.so")


def build(verbose: bool = False) -> str:
    """Compile the C/C++ restatement (oracle/libicg_oracle.so).  oracle/_ref (the reference's own sources) cannot
    be built for this reference: every hot-path source needs OpenCV / Ceres / Eigen headers absent from the image."""
    subprocess.run(["make", "-C", HERE] + ([] if verbose else ["-s"]), check=True)
    return LIB
