"""TEST / BASELINE INFRASTRUCTURE — build oracle/_build/librda_cpu_port.so with g++ (-O3, OpenMP)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'cpu_port', 'rda_cpu_port.cpp')
SO = os.path.join(HERE, '_build', 'librda_cpu_port.so')


def build(force=False):
    csrc = os.path.join(HERE, '..', 'rda_planner_b200', 'csrc')
    deps = [SRC, os.path.join(HERE, '..', 'include', 'rda_b200.h')] + [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if (not force) and os.path.exists(SO) and os.environ.get('RDA_B200_NO_BUILD') == '1':
        return SO
    if (not force) and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return SO
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(['g++', '-O3', '-std=c++17', '-fopenmp', '-shared', '-fPIC', '-o', SO, SRC])
    return SO


if __name__ == '__main__':
    print(build(force=True))
