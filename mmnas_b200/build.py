"""Builds libmmnas_b200.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

    python -m mmnas_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the repository snapshot."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'lib', 'libmmnas_b200.so')
SOURCES = ['elementwise.cu', 'gemm_simt.cu', 'gemm_tc.cu', 'layernorm.cu', 'attention.cu', 'attention_tc.cu', 'relbias.cu', 'relbias_mma.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '--use_fast_math=false', '-Xcompiler', '-fPIC', '-shared', '-Xptxas', '-v']


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, '..', 'include', 'mmnas_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    flags = [f for f in NVCC_FLAGS if f != '--use_fast_math=false']
    cmd = [nvcc] + flags + ['-o', LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed building libmmnas_b200.so')
    with open(os.path.join(os.path.dirname(LIB), 'ptxas.log'), 'w') as f:
        f.write(r.stdout + r.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
