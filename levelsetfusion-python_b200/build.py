"""Builds liblsf_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblsf_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",  # bit-exact with the reference's non-contracted float32 arithmetic
    "-Xcompiler", "-fPIC",
]
OBJ_DIR = os.path.join(HERE, "_obj")  # git-ignored object files: one per translation unit, compiled in parallel


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    built = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "lsf_b200.h")]
    return any(os.path.getmtime(d) > built for d in deps)


def build(force=False, verbose=False, defines=(), output=None):
    """defines / output: experiment builds (python build.py -DNAME=1 -o liblsf_exp.so; selected at run time with
    LSF_B200_LIBRARY=<path>); the default build has neither."""
    if defines or output:
        return _build_experiment(list(defines), output or os.path.join(HERE, "liblsf_b200_exp.so"))
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    env = dict(os.environ)
    env.pop("CXX", None)  # the image's CXX points at a compiler without OpenMP specs; let nvcc pick g++ from PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "lsf_b200.h")]
    newest_header = max(os.path.getmtime(h) for h in headers)

    def compile_unit(source):
        obj = os.path.join(OBJ_DIR, os.path.basename(source)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(source), newest_header):
            return obj, ""
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, source]
        result = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if result.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), result.stderr))
        return obj, result.stderr

    with ThreadPoolExecutor(max_workers=8) as pool:
        compiled = list(pool.map(compile_unit, sources()))
    if verbose:
        print("".join(log for _, log in compiled))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + [obj for obj, _ in compiled]
    result = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if result.returncode != 0:
        raise RuntimeError("nvcc link failed:\n%s\n%s" % (" ".join(cmd), result.stderr))
    return LIB


def _build_experiment(defines, output):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    env = dict(os.environ)
    env.pop("CXX", None)
    from concurrent.futures import ThreadPoolExecutor
    directory = os.path.join(OBJ_DIR, "exp_" + os.path.basename(output))
    os.makedirs(directory, exist_ok=True)

    def compile_unit(source):
        obj = os.path.join(directory, os.path.basename(source)[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + defines + ["-c", "-o", obj, source]
        result = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if result.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), result.stderr))
        return obj

    with ThreadPoolExecutor(max_workers=8) as pool:
        objects = list(pool.map(compile_unit, sources()))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", output] + objects
    result = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if result.returncode != 0:
        raise RuntimeError("nvcc link failed:\n%s" % result.stderr)
    return output


if __name__ == "__main__":
    import sys
    defines = [a for a in sys.argv[1:] if a.startswith("-D")]
    output = sys.argv[sys.argv.index("-o") + 1] if "-o" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defines,
                output=os.path.abspath(output) if output else None))
