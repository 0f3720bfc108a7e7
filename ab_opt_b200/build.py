"""Build libabopt_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m ab_opt_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, '_lib')
LIB_PATH = os.path.join(LIB_DIR, 'libabopt_b200.so')
SOURCES = ['api.cu', 'k_linear.cu', 'k_attn.cu', 'k_attn_tc.cu', 'k_pair.cu', 'k_step.cu', 'k_tc.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '--std=c++17',
              '-Xcompiler', '-fPIC', '-shared', '-Xptxas', '-v']


def _newest_source_mtime():
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    paths.append(os.path.join(os.path.dirname(HERE), 'include', 'abopt_b200.h'))
    return max(os.path.getmtime(p) for p in paths)


def build(force=False, verbose=False):
    """Compile every CUDA source into ab_opt_b200/_lib/libabopt_b200.so; returns the path."""
    os.makedirs(LIB_DIR, exist_ok=True)
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _newest_source_mtime():
        return LIB_PATH
    nvcc = os.environ.get('NVCC', 'nvcc')
    cmd = [nvcc] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ['-o', LIB_PATH]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    with open(os.path.join(LIB_DIR, 'build.log'), 'w') as f:
        f.write(' '.join(cmd) + '\n' + log)
    if proc.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + log[-4000:])
    if verbose:
        print(log)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
