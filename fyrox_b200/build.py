"""In-tree build of the native libraries (nvcc cross-compiles sm_100a without a GPU)."""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")


def _run(cmd, cwd):
    r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"{' '.join(cmd)} failed in {cwd}:\n{r.stdout}\n{r.stderr}")
    return r.stdout


def build_all(verbose: bool = False) -> None:
    """libfyrox_b200.so (CUDA, sm_100a) + libfyrox_scenegen.so (host) into fyrox_b200/lib/."""
    os.makedirs(os.path.join(_HERE, "lib"), exist_ok=True)
    out = _run(["make", "-C", CSRC, "all"], CSRC)
    if verbose:
        print(out)
