"""Build libtbk_b200.so (the C-ABI library) in-tree with nvcc for sm_100a.

    python -m pythtb_b200.build [--force] [--verbose]

The .so is written next to this file so that it travels with the repository
snapshot to the GPU box; it is git-ignored.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libtbk_b200.so")
SOURCES = ["tbk_api.cu", "tbk_solve.cu", "tbk_berry.cu"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC", "--fmad=true", "-Xptxas", "-v"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "tbk.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra_flags=None, out=None):
    """extra_flags / out: an A/B variant of the library (profiles/build_variant.py) — object files go to a scratch
    directory and the default in-tree library is left alone."""
    if extra_flags is None and out is None and not force and not needs_build():
        return SO
    objdir = CSRC
    so = SO
    if out is not None:
        import tempfile
        objdir = tempfile.mkdtemp(prefix="tbk_variant_")
        so = out
    # the three translation units are independent: compile them side by side
    from concurrent.futures import ThreadPoolExecutor

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags or []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return src, obj, res

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        done = list(pool.map(compile_one, SOURCES))
    objs = []
    logs = []
    for src, obj, res in done:
        logs.append(res.stdout)
        if res.returncode != 0:
            sys.stderr.write(res.stdout)
            raise RuntimeError("nvcc failed on " + src)
        objs.append(obj)
    cmd = [_nvcc(), "-shared", "-o", so] + objs + ["-lcudart"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("link failed")
    with open(os.path.join(objdir, "ptxas_info.log"), "w") as f:
        f.write("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    return so


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
