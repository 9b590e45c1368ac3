"""Build the CUDA shared library in-tree: facenet_b200/_lib/libfacenet_b200.so (sm_100a only).

    python -m facenet_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels with the repo snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / 'csrc'
OUT_DIR = PKG / '_lib'
LIB = OUT_DIR / 'libfacenet_b200.so'
SOURCES = ['fnb_api.cu', 'fnb_gram.cu', 'fnb_prepare.cu', 'fnb_select.cu', 'fnb_mine.cu', 'fnb_bce.cu', 'fnb_stage.cu', 'fnb_comm.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v']


def nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def _newest_source_mtime():
    files = list(CSRC.glob('*')) + [PKG.parent / 'include' / 'facenet_b200.h']
    return max(f.stat().st_mtime for f in files)


def needs_build():
    return not LIB.exists() or LIB.stat().st_mtime < _newest_source_mtime()


def build(force=False, verbose=True):
    if not force and not needs_build():
        return LIB
    OUT_DIR.mkdir(exist_ok=True)
    obj_dir = OUT_DIR / 'obj'
    obj_dir.mkdir(exist_ok=True)
    cc = nvcc()

    def compile_one(src):
        obj = obj_dir / (src.replace('.cu', '.o'))
        cmd = [cc] + NVCC_FLAGS + ['-c', str(CSRC / src), '-o', str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        # tracked ptxas -v log (registers, spills, barriers per kernel); compile times would make every rebuild a diff
        log = ''.join(ln for ln in (r.stdout + r.stderr).splitlines(keepends=True) if 'Compile time' not in ln)
        (obj_dir / (src + '.log')).write_text(log)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s' % (src, r.stdout + r.stderr))
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [cc, '-shared', '-o', str(LIB)] + [str(o) for o in objs] + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart_static', '-ldl', '-lrt', '-lpthread']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout + r.stderr)
    if verbose:
        print('built', LIB)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv)
