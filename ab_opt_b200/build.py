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
SOURCES = ['api.cu', 'k_linear.cu', 'k_attn.cu', 'k_attn_tc.cu', 'k_pair.cu', 'k_step.cu', 'k_tc.cu', 'k_tail_tc.cu', 'k_pair_embed.cu', 'k_res_embed.cu', 'k_post.cu', 'k_backward.cu']
ARCH_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a']
COMPILE_FLAGS = ARCH_FLAGS + ['-lineinfo', '-O3', '--std=c++17', '-Xcompiler', '-fPIC', '-Xptxas', '-v']
LINK_FLAGS = ARCH_FLAGS + ['-shared', '-Xcompiler', '-fPIC']


OBJ_DIR = os.path.join(LIB_DIR, 'obj')
HEADERS_GLOB = ('.cuh', '.h')


def _header_mtime():
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(HEADERS_GLOB)]
    paths.append(os.path.join(os.path.dirname(HERE), 'include', 'abopt_b200.h'))
    return max(os.path.getmtime(p) for p in paths)


def build(force=False, verbose=False, variant=None, defines=()):
    """Compile every CUDA source into ab_opt_b200/_lib/libabopt_b200.so; returns the path.
    One nvcc process per translation unit, run concurrently; objects are reused when neither the
    source nor any header is newer.
    variant / defines: an A/B build with extra -D macros into _lib/variants/<variant>/ (select it with ABOPT_LIB=<path>)."""
    obj_dir, lib_path = OBJ_DIR, LIB_PATH
    if variant:
        vdir = os.path.join(LIB_DIR, 'variants', variant)
        obj_dir, lib_path = os.path.join(vdir, 'obj'), os.path.join(vdir, 'libabopt_b200.so')
        force = True
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = os.environ.get('NVCC', 'nvcc')
    hdr = _header_mtime()
    jobs, objs, logs = [], [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(obj_dir, s[:-3] + '.o')
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr):
            continue
        cmd = [nvcc] + COMPILE_FLAGS + ['-D' + d for d in defines] + ['-c', src, '-o', obj]
        jobs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = None
    for cmd, proc in jobs:
        out, _ = proc.communicate()
        logs.append(' '.join(cmd) + '\n' + out)
        if proc.returncode != 0 and failed is None:
            failed = out
    if jobs:
        with open(os.path.join(LIB_DIR, 'build.log'), 'a' if not force else 'w') as f:
            f.write('\n'.join(logs))
    if failed is not None:
        raise RuntimeError('nvcc failed:\n' + failed[-4000:])
    if jobs or not os.path.exists(lib_path):
        cmd = [nvcc] + LINK_FLAGS + objs + ['-o', lib_path]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError('link failed:\n' + (proc.stdout + proc.stderr)[-4000:])
    if verbose:
        print('\n'.join(logs))
    return lib_path


if __name__ == '__main__':
    var = [a.split('=', 1)[1] for a in sys.argv if a.startswith('--variant=')]
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv, variant=var[0] if var else None,
                defines=[a[2:] for a in sys.argv if a.startswith('-D')]))
