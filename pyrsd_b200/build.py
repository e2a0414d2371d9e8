"""Build the native pieces in-tree.

  pyrsd_b200/libpyrsd_b200.so      the product: hand-written sm_100a CUDA + C-ABI (nvcc)
  tests/host_emul/libprsd_emul.so  test-only CPU emulation of the host logic (g++)
  oracle/_ref/libgcl_ref.so        test-only oracle: reference C++ compiled in place (only
                                   where /root/reference exists; otherwise the prebuilt file
                                   that travelled with the snapshot is used)

Run as ``python -m pyrsd_b200.build`` or through ``__graft_entry__.build()``.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libpyrsd_b200.so')

CU_SOURCES = ['context.cu', 'bilinear.cu', 'nodal.cu', 'linear.cu', 'spectra.cu', 'plan.cu', 'zeldovich.cu',
              'fftlog.cu', 'xilm.cu', 'discrete.cu', 'transfer.cu', 'exchange.cu', 'galaxy.cu', 'biasrel.cu', 'gp.cu', 'microbench.cu', 'capi.cu']
CXX_SOURCES = ['quad.cpp', 'fftlog_host.cpp']

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr'] + os.environ.get('PRSD_NVCC_EXTRA', '').split()


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd, cwd=None):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (' '.join(cmd), r.stdout))
    return r.stdout


def build_product(force=False, verbose=False):
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))]
    headers.append(os.path.join(ROOT, 'include', 'pyrsd_b200.h'))
    objs = []
    procs = []
    for src in CU_SOURCES + CXX_SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(objdir, src + '.o')
        objs.append(obj)
        if force or _newer(obj, [path] + headers):
            cmd = [nvcc] + NVCC_FLAGS + ['-x', 'cu', '-c', path, '-o', obj]
            if verbose:
                cmd.insert(1, '-Xptxas=-v')
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (' '.join(cmd), out))
        if verbose and out:
            print(out)
    if force or _newer(LIB, objs):
        _run([nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart'])
    return LIB


def build_emul(force=False):
    d = os.path.join(ROOT, 'tests', 'host_emul')
    lib = os.path.join(d, 'libprsd_emul.so')
    deps = [os.path.join(d, 'emul.cpp')] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)
                                            if f.endswith(('.h', '.cuh', '.cpp'))]
    if force or _newer(lib, deps):
        _run(['g++', '-O2', '-std=c++14', '-fopenmp', '-fPIC', '-shared', '-I/usr/local/cuda/include',
              '-x', 'c++', 'emul.cpp', os.path.join(CSRC, 'quad.cpp'), '-o', lib], cwd=d)
    return lib


def build_oracle():
    d = os.path.join(ROOT, 'oracle')
    if os.path.isdir('/root/reference/pyRSD/_gcl/cpp'):
        _run(['make', '-C', d, '-j8'])
    return os.path.join(d, '_ref', 'libgcl_ref.so')


def build_all(force=False, verbose=False):
    lib = build_product(force=force, verbose=verbose)
    build_emul(force=force)
    build_oracle()
    return lib


if __name__ == '__main__':
    print(build_all(force='--force' in sys.argv, verbose='-v' in sys.argv))
