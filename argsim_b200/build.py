"""Builds libargsim_b200.so (the product) and libargsim_b200_dev.so (development microbenchmarks, include/argsim_b200_dev.h)
in-tree with nvcc for sm_100a (no JIT cache: the .so files travel with the repo)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libargsim_b200.so')
DEV_LIB = os.path.join(HERE, 'libargsim_b200_dev.so')
SOURCES = ['kernels.cu', 'gemm_simt.cu', 'gemm_tc.cu', 'gru_generic.cu', 'gru_mma.cu', 'gru_tc.cu', 'plan.cpp', 'engine.cu', 'capi.cu']
DEV_SOURCES = ['xbench.cu', 'capi_dev.cu']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC',
         '--expt-relaxed-constexpr']


def _deps_mtime():
    return max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC)
               if f.endswith(('.cu', '.cuh', '.h', '.cpp'))) if os.path.isdir(CSRC) else 0


def needs_build():
    inc = os.path.join(HERE, '..', 'include')
    newest = max(_deps_mtime(), os.path.getmtime(os.path.join(inc, 'argsim_b200.h')), os.path.getmtime(os.path.join(inc, 'argsim_b200_dev.h')))
    return any((not os.path.exists(l)) or os.path.getmtime(l) < newest for l in (LIB, DEV_LIB))


def build(force=False, verbose=False):
    """compiles and links under an exclusive file lock (several ranks of one torchrun launch may get here at once);
    the library is linked to a temporary name and moved into place atomically."""
    import fcntl
    if not force and not needs_build():
        return LIB
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    with open(os.path.join(objdir, '.lock'), 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():   # another process built it while this one waited
                return LIB
            return _build_locked(objdir, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(objdir, verbose):

    def cc(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + '.o')
        cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-x', 'cu', '-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES) + len(DEV_SOURCES))) as ex:
        objs = list(ex.map(cc, SOURCES + DEV_SOURCES))
    for lib, lib_objs in ((DEV_LIB, objs[len(SOURCES):]), (LIB, objs[:len(SOURCES)])):
        tmp = lib + '.tmp.%d' % os.getpid()
        cmd = [NVCC, '-shared', '-o', tmp] + lib_objs + ['-Xcompiler', '-fPIC', '-ldl', '-cudart', 'static',
                                                          '-gencode', 'arch=compute_100a,code=sm_100a']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            if os.path.exists(tmp):
                os.unlink(tmp)
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
        os.replace(tmp, lib)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
