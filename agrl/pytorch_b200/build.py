"""Compile libagrl_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m agrl.pytorch_b200.build [--force] [--verbose]

One shared library, C ABI only (include/agrl_b200.h).  It links the CUDA runtime statically and has
no dependency on torch, so any FFI (ctypes here; cgo/JNI elsewhere) can load it.  The driver API
(cuTensorMapEncodeTiled for the TMA descriptors) is resolved at run time through
cudaGetDriverEntryPoint, hence no -lcuda at link time.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, 'libagrl_b200.so')
OBJ = os.path.join(HERE, 'csrc', '_obj')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-std=c++17', '-lineinfo',
    '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden',
    '--expt-relaxed-constexpr',
    '-Xfatbin', '-compress-all',             # the embedded cubins (with their -lineinfo tables) compressed: 5.8 -> 2.5 MB
    '-I', os.path.join(ROOT, 'include'),
]


def extra_flags():
    """AGRL_NVCC_EXTRA: extra nvcc flags for diagnostic builds (e.g. -DAGRL_TIMELINE, see tools/graph_timeline.py)."""
    return os.environ.get('AGRL_NVCC_EXTRA', '').split()


def nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: libagrl_b200 cannot be built (there is no CPU fallback)')
    return exe


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    return hs + [os.path.join(ROOT, 'include', 'agrl_b200.h'), os.path.abspath(__file__)]


def stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Build (if needed) and return the path of libagrl_b200.so."""
    srcs, hdrs = sources(), headers()
    if not force and not stale(LIB, srcs + hdrs):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    exe = nvcc()
    objs, procs = [], []
    for src in srcs:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        if force or stale(obj, [src] + hdrs):
            cmd = [exe] + NVCC_FLAGS + extra_flags() + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write('nvcc failed on %s:\n%s\n' % (src, out))
        elif verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError('nvcc failed')
    cmd = [exe, '-shared', '-o', LIB] + objs + ['-cudart', 'static', '-Xcompiler', '-fPIC']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
