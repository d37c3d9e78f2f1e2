"""Builds libslacken_gpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m slacken_b200.build [--force]

One translation unit per window width W (slk_inst.cu, -DSLK_W=n) plus the host/ABI unit and the sort unit, compiled
in parallel and linked into slacken_b200/libslacken_gpu.so with a static CUDA runtime.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
SO = os.path.join(HERE, "libslacken_gpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-ccbin", "/usr/bin/g++",
         "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v"]
WIDTHS = range(1, 9)
# tuning variants for A/B runs on the GPU box: SLK_VARIANT=name SLK_DEFS="-DSLK_CLS_MINB=5 -DSLK_LOOKUP_ILP=2"
# builds libslacken_gpu_<name>.so next to the default library; SLK_SO=libslacken_gpu_<name>.so selects it at load time
if os.environ.get("SLK_VARIANT"):
    FLAGS = FLAGS + os.environ.get("SLK_DEFS", "").split()
    SO = os.path.join(HERE, "libslacken_gpu_" + os.environ["SLK_VARIANT"] + ".so")
    OBJ = OBJ + "_" + os.environ["SLK_VARIANT"]


def _sources():
    deps = [os.path.join(CSRC, f) for f in ("slk_core.h", "slk_group.h", "slk_kernels.cuh", "slk_sort.h", "slk_host.h")]
    deps.append(os.path.join(HERE, "..", "include", "slacken_gpu.h"))
    units = [("slacken_gpu", os.path.join(CSRC, "slacken_gpu.cu"), []),
             ("slk_sort", os.path.join(CSRC, "slk_sort.cu"), []),
             ("slk_split", os.path.join(CSRC, "slk_split.cu"), [])]
    units += [(f"slk_inst_w{w}", os.path.join(CSRC, "slk_inst.cu"), [f"-DSLK_W={w}"]) for w in WIDTHS]
    return units, deps


def _stale(target, srcs):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


def _compile(unit, deps, force):
    name, src, extra = unit
    obj = os.path.join(OBJ, name + ".o")
    if not force and not _stale(obj, [src] + deps):
        return obj, ""
    cmd = [NVCC] + FLAGS + extra + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}\n{r.stderr}")
    with open(os.path.join(OBJ, name + ".ptxas.log"), "w") as f:
        f.write(r.stderr)
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    units, deps = _sources()
    # SLK_VARIANT_UNITS=slk_inst_w5[,..]: a variant recompiles only these units with its flags and links the default build's
    # objects for the rest (the tuning macros only reach the classify kernels)
    only = [u for u in os.environ.get("SLK_VARIANT_UNITS", "").split(",") if u] if os.environ.get("SLK_VARIANT") else []
    reuse = []
    if only:
        default_obj = os.path.join(HERE, "build")
        reuse = [os.path.join(default_obj, u[0] + ".o") for u in units if u[0] not in only]
        units = [u for u in units if u[0] in only]
    with ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 4)) as ex:
        results = list(ex.map(lambda u: _compile(u, deps, force), units))
    objs = [o for o, _ in results] + reuse
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    if force or _stale(SO, objs):
        cmd = [NVCC, "-shared", "-o", SO, "-ccbin", "/usr/bin/g++", "-cudart", "static"] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
