"""Compile the CUDA library in-tree with nvcc for sm_100a (no JIT cache, no torch headers)."""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libmuzero_b200.so')
SOURCES = ['api.cu', 'mcts.cu', 'mlp.cu', 'net.cu', 'conv.cu', 'selfplay.cu', 'replay.cu', 'train.cu', 'optim.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(verbose: bool = False, force: bool = False) -> str:
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(os.path.dirname(HERE), 'include', 'muzero_b200.h'))
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace('.cu', '.o'))
        if not force and _newer(o, [s] + headers):
            return o
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', s, '-o', o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError(f'nvcc failed on {src}')
        return o

    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or not _newer(OUT, objs):
        r = subprocess.run([nvcc, '-shared', '-o', OUT] + objs + ['-lcudart'], capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('link failed')
    return OUT


if __name__ == '__main__':
    print(build(verbose='-v' in sys.argv, force='-f' in sys.argv))
