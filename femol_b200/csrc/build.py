"""Build libfemol_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python femol_b200/csrc/build.py [--force] [--verbose]

The shared library is the C-ABI declared in include/femol_b200.h.  It has no torch / Python
dependency; the Python host loads it with ctypes (femol_b200/_lib.py).
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
SOURCES = ['plan.cu', 'elements.cu', 'assemble.cu', 'spmv.cu', 'pcg.cu', 'simp.cu', 'lobpcg.cu', 'gmg.cu', 'dist.cu', 'stencil.cu']
HEADERS = ['common.cuh', 'spmv.cuh', os.path.join(ROOT, 'include', 'femol_b200.h')]
OUT = os.path.join(PKG, 'libfemol_b200.so')
OBJ_DIR = os.path.join(HERE, 'build')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _digest():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS + [os.path.abspath(__file__)]:
        path = f if os.path.isabs(f) else os.path.join(HERE, f)
        with open(path, 'rb') as fh:
            h.update(fh.read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    stamp = os.path.join(OBJ_DIR, 'stamp')
    digest = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == digest:
        return OUT
    os.makedirs(OBJ_DIR, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(OBJ_DIR, src.replace('.cu', '.o'))
        cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(HERE, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write('---- %s ----\n%s\n' % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    cmd = [NVCC, '-shared', '-o', OUT] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-ldl']
    subprocess.check_call(cmd)
    with open(stamp, 'w') as fh:
        fh.write(digest)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
