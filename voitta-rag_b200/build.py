"""Build recipe for libvoitta_b200.so (nvcc, sm_100a only) — used by __graft_entry__.build()."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "csrc" / "vb_api.cu"
LIB = HERE / "libvoitta_b200.so"
FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-shared"]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; libvoitta_b200.so cannot be built")
    return p


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    srcs = list((HERE / "csrc").glob("*")) + [HERE.parent / "include" / "voitta_b200.h"]
    return any(s.stat().st_mtime > t for s in srcs)


LIB_DEBUG = HERE / "libvoitta_b200_dbg.so"


def build_debug() -> Path:
    """The same library with -DVB_DEBUG_BOUNDS: every VB_CHECK becomes a device-side assert (tools/gpu_sanitize.sh;
    select it with VB200_LIB=<path>).  compute-sanitizer is closed on the GPU pool; these are our own bounds checks."""
    cmd = [nvcc_path(), *FLAGS, "-DVB_DEBUG_BOUNDS", "-o", str(LIB_DEBUG), str(SRC)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{r.stdout}\n{r.stderr}")
    return LIB_DEBUG


def build_variant(name: str, *defines: str) -> Path:
    """A/B builds (tools/gpu_r2_m.sh): the library with extra -D flags as libvoitta_b200_<name>.so; select with VB200_LIB."""
    out = HERE / f"libvoitta_b200_{name}.so"
    r = subprocess.run([nvcc_path(), *FLAGS, *[f"-D{d}" for d in defines], "-o", str(out), str(SRC)], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{r.stdout}\n{r.stderr}")
    return out


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path(), *FLAGS, "-o", str(LIB), str(SRC)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
