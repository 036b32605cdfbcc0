"""Build libldot_sm100a.so in-tree with nvcc (sm_100a only).  `python -m lightningdot_b200.build [--force]`."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libldot_sm100a.so")
OBJ_DIR = os.path.join(CSRC, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newest_dep():
    t = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")):
                t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


HOSTEXT_SRC = os.path.join(HERE, "hostext", "pylists.c")
HOSTEXT_OUT = os.path.join(HERE, "_ldot_pyhost.so")


def build_hostext(force=False):
    """The small CPython helper behind DenseFlatIndexer.search_knn's id lists (hostext/pylists.c): host compiler only,
    no CUDA.  Optional - the indexer falls back to its numpy formulation of the same host-side loop without it."""
    import sysconfig
    if not force and os.path.exists(HOSTEXT_OUT) and os.path.getmtime(HOSTEXT_OUT) >= os.path.getmtime(HOSTEXT_SRC):
        return HOSTEXT_OUT
    cc = os.environ.get("CC", "gcc")
    cmd = [cc, "-O2", "-shared", "-fPIC", "-I" + sysconfig.get_paths()["include"], HOSTEXT_SRC, "-o", HOSTEXT_OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"host extension build failed:\n{r.stdout}\n{r.stderr}")
    return HOSTEXT_OUT


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ and link the shared library.  Returns the path of the .so."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    try:
        build_hostext(force)
    except Exception as exc:   # noqa: BLE001 - optional host-side helper: never block the CUDA build
        sys.stderr.write(f"warning: {exc}\n")
    dep_t = _newest_dep()
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= dep_t:
        return OUT

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= dep_t:
            return obj, ""
        cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, sources()))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
