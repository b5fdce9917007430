"""In-tree build of libvrsbs.so for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "vrsbs_api.cu")
OUT = os.path.join(_HERE, "libvrsbs.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def _sources():
    d = os.path.join(_HERE, "csrc")
    inc = os.path.join(os.path.dirname(_HERE), "include", "vrsbs.h")
    return [os.path.join(d, f) for f in sorted(os.listdir(d))] + [inc, os.path.abspath(__file__)]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force=False, verbose=False):
    """Compile csrc/vrsbs_api.cu -> libvrsbs.so. Returns the path."""
    if not force and not needs_build():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT + ".tmp", SRC]
    env = dict(os.environ)
    # the image exports CC/CXX=/opt/gcc/bin/*; nvcc's host compiler is chosen explicitly instead
    host = shutil.which("g++", path="/usr/bin") or shutil.which("g++")
    if host:
        cmd[1:1] = ["-ccbin", host]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed ({r.returncode}): {' '.join(cmd)}")
    os.replace(OUT + ".tmp", OUT)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
