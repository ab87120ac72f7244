"""Build libmbx.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m metabox_b200.build [--force] [--verbose]

One nvcc invocation per .cu (parallel), then a shared link.  The .so is git-ignored but travels to the
GPU box with the gpurun snapshot.  `-fmad=false`: a*b+c stays two IEEE ops (bit parity with the numpy
chains of the reference); dot products request fma() explicitly.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
BUILD = os.path.join(CSRC, 'build')
LIB = os.path.join(CSRC, 'libmbx.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
SOURCES = ['mbx_capi.cu', 'mbx_kernels.cu', 'mbx_rlepso.cu', 'mbx_lde.cu', 'mbx_deddqn.cu', 'mbx_deddqn_run.cu', 'mbx_gleet.cu', 'mbx_rlpso.cu', 'mbx_qlpso.cu', 'mbx_landscape.cu', 'mbx_gleet_policy.cu', 'mbx_deddqn_policy.cu', 'mbx_rlepso_policy.cu', 'mbx_pso_policies.cu', 'mbx_walk.cu']
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-fmad=false',
         '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def _deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh', '.h'))]
    out.append(os.path.join(HERE, '..', 'include', 'metabox_b200.h'))
    return out


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def _compile(src, verbose, extra=(), bdir=BUILD):
    obj = os.path.join(bdir, src.replace('.cu', '.o'))
    cmd = [NVCC] + FLAGS + list(extra) + ['-c', os.path.join(CSRC, src), '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(obj + '.log', 'w') as f:
        f.write(log)
    if r.returncode != 0:
        raise RuntimeError(f'nvcc failed for {src}:\n{log[-6000:]}')
    if verbose:
        print(log)
    return obj


def build(force=False, verbose=False, profile=False, variant=None, defines=()):
    """profile=True builds the phase-instrumented variant libmbx_prof.so (-DMBX_PROFILE; measurement only, selected with
    the MBX_LIB environment variable -- the product library is never instrumented).  variant='name' with defines=['-DMBX_FDR_UNROLL=4', ...]
    builds libmbx_<name>.so with those compile-time tuning knobs for A/B measurements (scripts/variant_sweep.sh); the product library is
    always the default configuration."""
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    tag = 'prof' if profile else variant
    lib = LIB.replace('libmbx.so', f'libmbx_{tag}.so') if tag else LIB
    bdir = BUILD + f'_{tag}' if tag else BUILD
    if not tag and not force and not needs_build():
        return LIB
    os.makedirs(bdir, exist_ok=True)
    extra = (['-DMBX_PROFILE=1'] if profile else []) + list(defines)
    with cf.ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose, extra, bdir), srcs))
    cmd = [NVCC, '-shared', '-o', lib] + objs + ['-lcudart']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout + r.stderr)
    return lib


if __name__ == '__main__':
    var = sys.argv[sys.argv.index('--variant') + 1] if '--variant' in sys.argv else None
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv, profile='--profile' in sys.argv, variant=var,
                defines=[a for a in sys.argv[1:] if a.startswith('-D')]))
