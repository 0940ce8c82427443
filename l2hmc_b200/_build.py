"""Builds `l2hmc_b200/libl2b.so` (the C-ABI CUDA library, include/l2b.h) in-tree
with nvcc for sm_100a.  No torch headers are involved: the library's interface
is plain pointers and sizes, and Python talks to it through ctypes
(`l2hmc_b200/_lib.py`)."""
from __future__ import annotations

import fcntl
import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / 'csrc'
LIB = PKG / 'libl2b.so'
SOURCES = ['l2b_capi.cu', 'l2b_su3.cu', 'l2b_u1.cu', 'l2b_vnet.cu', 'l2b_gemm.cu', 'l2b_conv.cu']
NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3',
    '-std=c++17', '-Xcompiler', '-fPIC',
]


def _nvcc() -> str:
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: cannot build libl2b.so')
    return exe


HEADER = PKG.parent / 'include' / 'l2b.h'


def abi_hash() -> int:
    """first 32 bits of SHA-256(include/l2b.h): what `l2b_abi_hash()` of a matching binary returns"""
    return int.from_bytes(hashlib.sha256(HEADER.read_bytes()).digest()[:4], 'big')


def source_hash() -> int:
    """first 32 bits of SHA-256 over the header and every file under csrc/ (names and bytes):
    what `l2b_source_hash()` of a binary built from this tree returns"""
    h = hashlib.sha256()
    for f in [HEADER] + sorted(CSRC.glob('*.cu*')):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return int.from_bytes(h.digest()[:4], 'big')


def _stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob('*.cu')) + list(CSRC.glob('*.cuh')) + [PKG.parent / 'include' / 'l2b.h']
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not _stale():
        return LIB
    nvcc = _nvcc()
    objdir = PKG / 'build'
    objdir.mkdir(exist_ok=True)
    # one builder at a time (torchrun ranks importing the package concurrently); whoever gets the lock
    # second finds a fresh library and returns
    with open(objdir / '.lock', 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():
                return LIB
            return _build_locked(nvcc, objdir, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(nvcc: str, objdir: Path, verbose: bool) -> Path:
    defs = [f'-DL2B_ABI_HASH={abi_hash()}u', f'-DL2B_SRC_HASH={source_hash()}u']

    def compile_one(src: str) -> Path:
        obj = objdir / (src + '.o')
        cmd = [nvcc, *NVCC_FLAGS, *(defs if src == 'l2b_capi.cu' else []), '-c', str(CSRC / src), '-o', str(obj)]
        if verbose:
            cmd[1:1] = ['-Xptxas', '-v']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{r.stdout}\n{r.stderr}')
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB.with_suffix('.so.tmp')
    cmd = [nvcc, '-shared', '-o', str(tmp), *map(str, objs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    os.replace(tmp, LIB)        # atomic: a concurrent dlopen sees the old or the new file, never half of one
    return LIB


if __name__ == '__main__':
    import sys
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
