"""Build libmome.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m exploremultimodal_b200.build_ext [--force]

One translation unit per csrc/*.cu, compiled in parallel, linked into
exploremultimodal_b200/libmome.so. The CUDA runtime is linked statically and the driver entry
point for TMA descriptors is resolved at run time, so the library loads on a machine without a GPU
(the ABI test runs there).
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
BUILD = os.path.join(HERE, 'csrc', 'build')
LIB = os.path.join(HERE, 'libmome.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v']


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


HASH = LIB + '.srchash'


def _sources_hash():
    """sha256 over every csrc/*.cu, csrc/*.cuh, include/mome.h and the compiler flags."""
    import hashlib
    h = hashlib.sha256(' '.join(FLAGS).encode())
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh')))
    files.append(os.path.join(os.path.dirname(HERE), 'include', 'mome.h'))
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, 'rb') as fh:
            h.update(fh.read())
    return h.hexdigest()


def needs_build():
    """True unless libmome.so exists and was built from exactly the present sources (content hash, not mtimes:
    the library travels to the GPU box inside a snapshot whose file times mean nothing)."""
    if not os.path.exists(LIB) or not os.path.exists(HASH):
        return True
    with open(HASH) as f:
        return f.read().strip() != _sources_hash()


def _compile(src):
    obj = os.path.join(BUILD, src[:-3] + '.o')
    cmd = [NVCC, *FLAGS, '-c', os.path.join(CSRC, src), '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(obj + '.log', 'w') as f:
        f.write(' '.join(cmd) + '\n' + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f'nvcc failed on {src}:\n{r.stdout}\n{r.stderr}')
    return obj


def build(force=False, verbose=True):
    if not force and not needs_build():
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    srcs = _sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(_compile, srcs))
    cmd = [NVCC, '-shared', '-o', LIB, *objs, '-gencode', 'arch=compute_100a,code=sm_100a', '-cudart', 'static']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    with open(HASH, 'w') as f:
        f.write(_sources_hash() + '\n')
    if verbose:
        print(f'built {LIB} from {len(srcs)} sources')
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv)
