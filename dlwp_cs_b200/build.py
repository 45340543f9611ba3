"""Build libdlwpcs.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m dlwp_cs_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libdlwpcs.so')
SOURCES = ['cs_api.cu', 'cs_fp32.cu', 'cs_tc.cu', 'cs_tc_rs.cu', 'cs_wgrad_tc.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-O3', '-I', os.path.join(ROOT, 'include'), '-I', CSRC]
# --use_fast_math (flush-to-zero, approximate division / square root) only for the two bf16 tensor-core translation
# units, whose arithmetic outside the tensor core is fmaf / fminf / fmaxf on values that are rounded to bf16 anyway.  The
# float32 parity kernels, the Adam update (m / (sqrt(v) + eps)), the loss and the insolation are compiled IEEE.
FAST_MATH = {'cs_tc.cu', 'cs_tc_rs.cu', 'cs_wgrad_tc.cu'}


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, 'include', 'dlwpcs.h'), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile the shared library if it is missing or older than its sources; returns its path."""
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + (['--use_fast_math'] if src in FAST_MATH else []) + \
            (['-Xptxas', '-v'] if verbose else []) + (['-DRS_DEBUG_WAITS'] if os.environ.get('DLWPCS_RS_DEBUG') else []) + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError('nvcc failed on %s' % src)
    cmd = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB_PATH] + objs + ['-lcudart']
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
