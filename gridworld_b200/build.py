"""Builds gridworld_b200/csrc/libiglu_b200.so (sm_100a) in-tree with nvcc.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
SO = os.path.join(CSRC, 'libiglu_b200.so')
SOURCES = ['iglu_b200.cu']
HEADERS = ['iglu_physics.cuh', 'iglu_reward.cuh', 'iglu_render.cuh', os.path.join('..', '..', 'include', 'iglu_b200.h')]

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '-std=c++17',
    '-fmad=false',            # fp64 parity: the reference never fuses a*b+c
    '--shared', '-Xcompiler', '-fPIC',
    '-Xptxas', '-v',
]


def _nvcc():
    for cand in (shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: the CUDA library cannot be built')


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


MB_SO = os.path.join(CSRC, 'libiglu_microbench.so')


def build_microbench(force=False):
    """Integer-pipe peak microbenchmark (POPC / LOP3 / IADD3 / IMAD issue rates; bench.py's reward roofline)."""
    src = os.path.join(CSRC, 'iglu_microbench.cu')
    if not force and os.path.exists(MB_SO) and os.path.getmtime(MB_SO) >= os.path.getmtime(src):
        return MB_SO
    cmd = [_nvcc(), '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '--shared',
           '-Xcompiler', '-fPIC', '-o', MB_SO, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        print(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed building libiglu_microbench.so')
    return MB_SO


def build_variant(name, defines, verbose=False):
    """Tuning aid: builds csrc/libiglu_b200_<name>.so with extra -D defines (select it with the
    IGLU_B200_SO environment variable)."""
    so = os.path.join(CSRC, f'libiglu_b200_{name}.so')
    cmd = [_nvcc()] + NVCC_FLAGS + [f'-D{d}' for d in defines] + ['-o', so] + [os.path.join(CSRC, f) for f in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f'nvcc failed building {so}')
    return so


def build_library(force=False, verbose=False):
    """Compile the CUDA library if it is missing or stale; returns the path of the .so."""
    override = os.environ.get('IGLU_B200_SO')
    if override:
        return override
    if not force and not needs_build():
        return SO
    # Several ranks may get here at once (bench.py under torchrun): one builds, under a file lock, into a
    # temporary name that is moved into place atomically; the others wait and then find a fresh library.
    import fcntl
    with open(os.path.join(CSRC, '.build.lock'), 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                return SO
            tmp = f'{SO}.{os.getpid()}.tmp'
            cmd = [_nvcc()] + NVCC_FLAGS + ['-o', tmp] + [os.path.join(CSRC, f) for f in SOURCES]
            res = subprocess.run(cmd, capture_output=True, text=True)
            log = res.stdout + res.stderr
            with open(os.path.join(CSRC, 'build.log'), 'w') as f:
                f.write(' '.join(cmd) + '\n' + log)
            if verbose or res.returncode != 0:
                print(log)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError('nvcc failed building libiglu_b200.so (see gridworld_b200/csrc/build.log)')
            os.replace(tmp, SO)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return SO


if __name__ == '__main__':
    print(build_library(force=True, verbose=True))
