"""Build libtfnas_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension)."""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libtfnas_b200.so')
SOURCES = ['api.cu', 'body.cu', 'stem.cu', 'head.cu', 'optim.cu', 'fwd.cu', 'bwd.cu', 'stage.cu', 'umma_selftest.cu', 'umma_pw.cu', 'umma_ws.cu', 'dws.cu']
HEADERS = ['common.cuh', 'kernels.h', 'api_internal.h', 'pw.cuh', 'umma.cuh', os.path.join('..', '..', 'include', 'tfnas_b200.h')]
NVCC_FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-gencode', 'arch=compute_100a,code=sm_100a',
              '-Xcompiler', '-fPIC', '-DTFNAS_NO_FAST_MATH',
              '-Xptxas', '-v' if os.environ.get('TFNAS_PTXAS_V') else '-O3', '-Xptxas', '-warn-spills']


def _nvcc():
    for c in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if c and os.path.exists(c):
            return c
    raise RuntimeError('nvcc not found')


def _digest():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), 'rb') as fh:
            h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


TRACE_LIB = os.path.join(LIBDIR, 'libtfnas_b200_trace.so')


def build(force=False, verbose=True, trace=False):
    """trace=True builds the debug variant with the in-kernel phase trace (-DUM_TRACE) next to the product library."""
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, 'libtfnas_b200%s.digest' % ('_trace' if trace else ''))
    dg = _digest()
    lib = TRACE_LIB if trace else LIB
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read().strip() == dg:
        return lib
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace('.cu', '_trace.o' if trace else '.o'))
        cmd = [nvcc] + NVCC_FLAGS + (['-DUM_TRACE'] if trace else []) + ['-c', os.path.join(CSRC, src), '-o', obj]
        if verbose:
            print(' '.join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            sys.stderr.write(out.decode(errors='replace'))
        if p.returncode != 0:
            raise RuntimeError('nvcc failed on %s:\n%s' % (src, out.decode(errors='replace')))
    cmd = [nvcc, '-shared', '-o', lib] + objs + ['-lcudart']
    subprocess.check_call(cmd)
    with open(stamp, 'w') as f:
        f.write(dg)
    return lib


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, trace='--trace' in sys.argv))
