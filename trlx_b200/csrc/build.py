"""In-tree build of the sm_100a extension → ``trlx_b200/_C.so``.

``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` for every ``.cu`` (plain C ABI, no torch
headers → seconds per file), ``g++`` for the pybind/torch bindings, one link step.  The result is
kept in-tree (git-ignored) so it travels with a repo snapshot to GPU boxes.  A content hash of the
sources makes rebuilds no-ops.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
PKG = HERE.parent
BUILD = HERE / "build"
TARGET = PKG / "_C.so"
CUDA_HOME = Path(os.environ.get("CUDA_HOME", "/usr/local/cuda"))
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _sources():
    cu = sorted(HERE.glob("*.cu"))
    cpp = sorted(HERE.glob("*.cpp"))
    hdr = sorted(HERE.glob("*.cuh")) + sorted(HERE.glob("*.h"))
    return cu, cpp, hdr


def _digest(files) -> str:
    h = hashlib.sha256()
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    import torch

    h.update(torch.__version__.encode())
    return h.hexdigest()


def _run(cmd, verbose):
    if verbose:
        print(" ".join(str(c) for c in cmd), flush=True)
    res = subprocess.run([str(c) for c in cmd], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"build step failed:\n{' '.join(str(c) for c in cmd)}\n{res.stdout}\n{res.stderr}")
    return res.stdout + res.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    import torch
    from torch.utils import cpp_extension

    cu, cpp, hdr = _sources()
    stamp = PKG / "_C.so.stamp"
    digest = _digest(cu + cpp + hdr)
    if not force and TARGET.exists() and stamp.exists() and stamp.read_text() == digest:
        return TARGET
    BUILD.mkdir(exist_ok=True)
    nvcc = CUDA_HOME / "bin" / "nvcc"
    if not nvcc.exists():
        found = shutil.which("nvcc")
        if not found:
            raise RuntimeError("nvcc not found")
        nvcc = Path(found)

    cu_flags = ARCH_FLAGS + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
                             "-Xptxas", "-v", f"-I{HERE}"]
    inc = [f"-I{p}" for p in cpp_extension.include_paths(device_type="cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cxx_flags = ["-O2", "-std=c++17", "-fPIC", f"-D_GLIBCXX_USE_CXX11_ABI={abi}", "-DTORCH_EXTENSION_NAME=_C",
                 "-DTORCH_API_INCLUDE_EXTENSION_H", "-w"] + inc

    jobs = []
    objs = []
    for src in cu:
        obj = BUILD / (src.stem + ".o")
        objs.append(obj)
        jobs.append([nvcc, *cu_flags, "-c", src, "-o", obj])
    for src in cpp:
        obj = BUILD / (src.stem + ".o")
        objs.append(obj)
        jobs.append(["g++", *cxx_flags, "-c", src, "-o", obj])
    with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
        logs = list(pool.map(lambda c: _run(c, verbose), jobs))
    (BUILD / "ptxas.log").write_text("\n".join(logs))

    lib_dirs = cpp_extension.library_paths(device_type="cuda")
    tmp_target = TARGET.with_suffix(".so.tmp")  # linked aside and renamed: a concurrent snapshot never sees a half-written library
    link = ["g++", "-shared", "-o", tmp_target, *objs]
    for d in lib_dirs:
        link += [f"-L{d}", f"-Wl,-rpath,{d}"]
    link += ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"]
    _run(link, verbose)
    os.replace(tmp_target, TARGET)
    stamp.write_text(digest)
    return TARGET


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose=True)
    print(f"built {path}")
