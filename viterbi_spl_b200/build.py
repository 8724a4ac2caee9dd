"""Builds libvit_b200.so (the CUDA decoder behind include/vit_b200.h) in-tree with nvcc for sm_100a.

    python -m viterbi_spl_b200.build [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with the repository snapshot to the GPU box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libvit_b200.so')
SOURCES = ['vit_api.cu', 'vit_backpointer.cu', 'vit_cluster.cu', 'vit_tmem.cu', 'vit_fb.cu', 'vit_emis.cu', 'vit_banded.cu', 'vit_banded_wide.cu', 'vit_fb_tc.cu', 'vit_stream.cu', 'vit_fb_banded.cu', 'vit_fb_conv.cu']
HEADERS = ['vit_common.cuh', 'vit_tmem.cuh', os.path.join('..', '..', 'include', 'vit_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '--compiler-options', '-fPIC', '-shared', '-Xptxas', '-v']


# the sources a kernel's machine code depends on: their hash is the "build id" that ties an ncu capture under profiles/
# (DRAM traffic per launch) to the library bench.py is timing -- a capture of an older kernel is refused, not reported
KERNEL_SOURCES = {
    'tmem_forward_kernel': ['vit_tmem.cu', 'vit_tmem.cuh', 'vit_common.cuh'],
    'stream_forward_kernel': ['vit_stream.cu', 'vit_tmem.cuh', 'vit_common.cuh'],
    'cluster_forward_kernel': ['vit_cluster.cu', 'vit_common.cuh'],
    'bp_forward_kernel': ['vit_backpointer.cu', 'vit_common.cuh'],
    'banded_forward_kernel': ['vit_banded.cu', 'vit_tmem.cuh', 'vit_common.cuh'],
    'wide_forward_kernel': ['vit_banded_wide.cu', 'vit_tmem.cuh', 'vit_common.cuh'],
    'fb_tc_pass_kernel': ['vit_fb_tc.cu', 'vit_common.cuh'],
    'fb_pass_kernel': ['vit_fb.cu', 'vit_tmem.cuh', 'vit_common.cuh'],
    'fb_banded_pass_kernel': ['vit_fb_banded.cu', 'vit_tmem.cuh', 'vit_common.cuh'],
    'fb_conv_pass_kernel': ['vit_fb_conv.cu', 'vit_common.cuh'],
    'emissions_reg_kernel': ['vit_emis.cu', 'vit_common.cuh'],
}


def kernel_build_id(kernel):
    """16 hex digits identifying the build of one kernel: sha256 over its source files and the nvcc flags."""
    import hashlib
    h = hashlib.sha256(' '.join(NVCC_FLAGS).encode())
    for name in KERNEL_SOURCES[kernel]:
        with open(os.path.join(CSRC, name), 'rb') as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: the CUDA decoder cannot be built')
    return exe


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source (one nvcc per file, in parallel) and link them into one shared library; returns its
    path."""
    if not force and not stale():
        return LIB
    import tempfile
    from concurrent.futures import ThreadPoolExecutor
    env = dict(os.environ)
    env.pop('CC', None)   # the image exports a wrapper CC that nvcc must not pick up as host compiler
    compile_flags = [f for f in NVCC_FLAGS if f != '-shared']
    logs = []
    with tempfile.TemporaryDirectory(prefix='vit_build_') as tmp:
        def compile_one(src):
            obj = os.path.join(tmp, src.replace('.cu', '.o'))
            cmd = [nvcc()] + compile_flags + ['-c', '-o', obj, os.path.join(CSRC, src)]
            res = subprocess.run(cmd, capture_output=True, text=True, env=env)
            if res.returncode != 0:
                raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + res.stdout + res.stderr)
            return obj, res.stdout + res.stderr
        with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
            results = list(pool.map(compile_one, SOURCES))
        objs = [o for o, _ in results]
        logs = [l for _, l in results]
        cmd = [nvcc(), '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '--compiler-options', '-fPIC', '-o', LIB] + objs
        res = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if res.returncode != 0:
            raise RuntimeError('nvcc link failed:\n' + ' '.join(cmd) + '\n' + res.stdout + res.stderr)
        logs.append(res.stdout + res.stderr)
    if verbose:
        print(''.join(logs))
    with open(os.path.join(HERE, 'build_ptxas.log'), 'w') as fh:
        fh.write(''.join(logs))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
