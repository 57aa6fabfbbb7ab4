"""Build libgiraffe_b200.so (sm_100a CUDA kernels + C-ABI + host index builder) in-tree.

The shared object is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "vg_b200" / "csrc"
LIB = ROOT / "vg_b200" / "libgiraffe_b200.so"
STAMP = ROOT / "vg_b200" / ".build_stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-sign-compare,-Wno-unused-function",
    "--expt-relaxed-constexpr",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cpp")))


def _digest() -> str:
    h = hashlib.sha256()
    h.update(os.environ.get("GB_TILE_DEBUG", "").encode())
    for p in sorted(list(CSRC.glob("*")) + [ROOT / "include" / "giraffe_b200.h", Path(__file__)]):
        if p.is_file():
            h.update(p.name.encode())
            h.update(p.read_bytes())
    return h.hexdigest()


def _header_digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [ROOT / "include" / "giraffe_b200.h", Path(__file__)]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """One object per translation unit (compiled in parallel, reused when neither the unit nor any header changed),
    then one link."""
    digest = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text() == digest:
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    objdir = ROOT / "build" / "obj"
    objdir.mkdir(parents=True, exist_ok=True)
    hd = _header_digest()
    base = [_nvcc(), *NVCC_FLAGS, "-I", str(ROOT / "include"), "-I", str(CSRC)]
    if os.path.exists("/usr/bin/g++"):
        base += ["-ccbin", "/usr/bin/g++"]
    if verbose:
        base += ["-Xptxas", "-v"]
    if os.environ.get("GB_TILE_DEBUG"):
        base += ["-DGB_TILE_DEBUG"]

    def compile_one(src: Path):
        obj = objdir / (src.name + ".o")
        stamp = objdir / (src.name + ".stamp")
        want = hashlib.sha256(src.read_bytes() + hd.encode() + os.environ.get("GB_TILE_DEBUG", "").encode()).hexdigest()
        if not force and not verbose and obj.exists() and stamp.exists() and stamp.read_text() == want:
            return obj, None
        res = subprocess.run(base + ["-c", "-o", str(obj), str(src)], capture_output=True, text=True)
        if res.returncode != 0:
            return obj, res.stdout + res.stderr
        if verbose:
            sys.stderr.write(res.stdout + res.stderr)
        stamp.write_text(want)
        return obj, None

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, _sources()))
    errors = [e for _, e in results if e]
    if errors:
        sys.stderr.write("\n".join(errors))
        raise RuntimeError("nvcc failed")
    tmp = LIB.with_suffix(".so.tmp")          # link to a scratch name, then rename: a reader never sees a torn library
    link = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-shared", "-o", str(tmp)]
    if os.path.exists("/usr/bin/g++"):
        link += ["-ccbin", "/usr/bin/g++"]
    link += [str(o) for o, _ in results]
    link += ["-lz"]                           # libz.so.1 is part of the image (the distribution's libz.a is not PIC)
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("link failed")
    os.replace(tmp, LIB)
    STAMP.write_text(digest)
    return LIB


def build_oracle(arch: str = "") -> Path:
    """Compile oracle/liboracle.so (test infrastructure; never loaded by the product)."""
    odir = ROOT / "oracle"
    env = dict(os.environ)
    res = subprocess.run(["make", "-C", str(odir), f"ARCH={arch}"], capture_output=True, text=True, env=env)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("oracle build failed")
    return odir / "liboracle.so"


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_oracle())
