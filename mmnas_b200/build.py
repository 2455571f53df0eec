"""Builds libmmnas_b200.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

    python -m mmnas_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the repository snapshot."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'lib', 'libmmnas_b200.so')
SOURCES = ['elementwise.cu', 'gemm_simt.cu', 'gemm_tc.cu', 'layernorm.cu', 'attention.cu', 'attention_tc.cu', 'relbias.cu', 'relbias_mma.cu', 'blocks.cu', 'gemm_ln.cu', 'lstm.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '--use_fast_math=false', '-Xcompiler', '-fPIC', '-shared', '-Xptxas', '-v']


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, '..', 'include', 'mmnas_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu for sm_100a (one nvcc process per source, in parallel) and link the shared library."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    objdir = os.path.join(os.path.dirname(LIB), 'obj')
    os.makedirs(objdir, exist_ok=True)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    flags = [f for f in NVCC_FLAGS if f not in ('--use_fast_math=false', '-shared')]

    def compile_one(src):
        obj = os.path.join(objdir, src.replace('.cu', '.o'))
        r = subprocess.run([nvcc] + flags + ['-c', '-o', obj, os.path.join(CSRC, src)], capture_output=True, text=True)
        return src, obj, r

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    log = ''.join('== %s\n%s%s' % (src, r.stdout, r.stderr) for src, _, r in results)
    failed = [src for src, _, r in results if r.returncode != 0]
    if verbose or failed:
        sys.stderr.write(log)
    if failed:
        raise RuntimeError('nvcc failed on %s' % ', '.join(failed))
    r = subprocess.run([nvcc, '-shared', '-o', LIB] + [obj for _, obj, _ in results], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('linking libmmnas_b200.so failed')
    with open(os.path.join(os.path.dirname(LIB), 'ptxas.log'), 'w') as f:
        f.write(log)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
