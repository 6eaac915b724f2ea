"""Build libdae_sm100.so (the C-ABI library of sm_100a kernels) in-tree with nvcc.

    python -m dae_rnn_news_recommendation_b200.build [--force]

Objects are cached under build/ keyed by source mtime; the .so lands next to this file so it travels with the
repo snapshot to the GPU box (it is git-ignored).  nvcc cross-compiles sm_100a without a GPU.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / 'csrc'
OUT = PKG / 'libdae_sm100.so'
OBJ_DIR = ROOT / 'build' / 'obj'

NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _sources():
    return sorted(CSRC.glob('*.cu'))


def _headers_mtime():
    hs = list(CSRC.glob('*.cuh')) + list((ROOT / 'include').glob('*.h'))
    return max(h.stat().st_mtime for h in hs)


def build(force=False, verbose=False):
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    hm = _headers_mtime()
    srcs = _sources()
    jobs = []
    for s in srcs:
        o = OBJ_DIR / (s.stem + '.o')
        if force or not o.exists() or o.stat().st_mtime < max(s.stat().st_mtime, hm):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [NVCC, *FLAGS, '-c', str(s), '-o', str(o)]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
        r = subprocess.run(cmd, capture_output=True, text=True)
        return s, r

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for s, r in ex.map(compile_one, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError('nvcc failed on %s' % s)
    objs = [OBJ_DIR / (s.stem + '.o') for s in srcs]
    if force or jobs or not OUT.exists():
        cmd = [NVCC, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', str(OUT), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('link failed')
    return OUT


if __name__ == '__main__':
    p = build(force='--force' in sys.argv, verbose='-v' in sys.argv)
    print(p)
