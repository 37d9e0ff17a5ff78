"""Build the in-tree CUDA library (sm_100a only) with nvcc.

    python -m humaniflow_b200.build

Produces humaniflow_b200/lib/libhumaniflow_b200.so next to the sources (git-ignored; it travels to the GPU
box with the repo snapshot).  cudart is linked statically; libcuda is not linked (the one driver entry
point needed, cuTensorMapEncodeTiled, is resolved at run time), so the library loads on a GPU-less host.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libhumaniflow_b200.so')
SOURCES = ['api.cu', 'lbs.cu', 'sampling.cu', 'metrics.cu', 'proxy_rep.cu', 'flow.cu', 'heads.cu', 'encoder.cu']
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']


def nvcc_path():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found')


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, '..', 'include', 'humaniflow_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = nvcc_path()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace('.cu', '.o'))
        cmd = [nvcc, *ARCH, '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC', '-c', os.path.join(CSRC, src), '-o', obj]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError('nvcc failed on %s:\n%s' % (src, out))
        if verbose:
            print(out)
    cmd = [nvcc, *ARCH, '-shared', '-Xcompiler', '-fPIC', '-cudart', 'static', '-o', LIB_PATH, *objs]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc link failed:\n%s' % r.stdout)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
