"""In-tree build of librda_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRCS = [os.path.join(HERE, 'csrc', 'rda_kernels.cu'), os.path.join(HERE, 'csrc', 'rda_frontend.cu')]
OUT = os.path.join(HERE, 'librda_b200.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC']


def needs_build():
    if not os.path.exists(OUT):
        return True
    if os.environ.get('RDA_B200_NO_BUILD') == '1':      # use the shipped library as is (GPU job scripts)
        return False
    deps = [os.path.join(HERE, 'csrc', f) for f in os.listdir(os.path.join(HERE, 'csrc'))]
    deps.append(os.path.join(HERE, '..', 'include', 'rda_b200.h'))
    return any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    out = os.environ.get('RDA_B200_BUILD_OUT', OUT)      # tuning experiments: variant libraries next to the default one
    extra = os.environ.get('RDA_B200_NVCC_EXTRA', '').split()
    cmd = [nvcc] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + ['-o', out] + SRCS
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
