"""Build the in-tree CUDA shared library (sm_100a only) with nvcc.  `python -m panonerf_b200.build`"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpanonerf_b200.so")
SOURCES = ["api.cu", "rays.cu", "render.cu", "shade.cu", "gemm_simt.cu", "gemm_tc.cu", "wgrad_batch.cu", "mlp_fused.cu", "image.cu", "variants.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         # un-fused fp32 arithmetic so element-wise kernels round like the reference's PyTorch ops; the GEMM inner
         # loops use explicit fmaf()
         "-fmad=false", "-Xcompiler", "-fPIC"]
if os.environ.get("PNB_PTXAS_V"):
    FLAGS += ["-Xptxas", "-v"]
if os.environ.get("PNB_EXTRA_NVCC_FLAGS"):          # e.g. -DPNB_MBAR_WATCHDOG: stuck barriers trap instead of hanging
    FLAGS += os.environ["PNB_EXTRA_NVCC_FLAGS"].split()


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "panonerf_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src} ---\n{out}\n")
        failed |= pr.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-cudart", "static"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
