"""In-tree build of libssd_gpu.so (CUDA kernels for sm_100a + C ABI + host classes).

    python -m stair_step_detector_b200.build

nvcc cross-compiles without a GPU. -fmad=false / -ffp-contract=off: the reference build never contracts
a*b+c and bit-exact labels need identical roundings.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libssd_gpu.so")
SCENE_LIB = os.path.join(LIB_DIR, "libssd_scene.so")  # synthetic input source (include/ssd_scene.h): not the product
SCENE_SOURCES = ["scene/ssd_scene.cu"]

SOURCES = [
    "ssd_gpu.cu",
    "ssd_host.cpp",
    "host/transformation.cpp",
    "host/stairs.cpp",
    "host/pointcloud.cpp",
    "host/segmentation.cpp",
    "host/quadrilateralTest.cpp",
    "host/defaultContext.cpp",
    "host/calibrationTriangle.cpp",
    "host/geometricCalibration.cpp",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=default",
    "-Xptxas", "-v",
    "-shared", "-cudart", "static",
]


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def _build_one(out, sources, force, verbose, log=None):
    srcs = [os.path.join(CSRC, s) for s in sources if os.path.exists(os.path.join(CSRC, s))]
    deps = list(srcs) + [os.path.join(HERE, "..", "include", "ssd_gpu.h"), os.path.join(HERE, "..", "include", "ssd_scene.h"), __file__]
    for root, _, files in os.walk(CSRC):
        deps += [os.path.join(root, f) for f in files if f.endswith((".h", ".cuh"))]
    if not force and not _stale(out, deps):
        return out
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", out] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building " + os.path.basename(out))
    if log:
        with open(os.path.join(LIB_DIR, log), "w") as f:
            f.write(r.stderr)
    return out


def build(force=False, verbose=False):
    """libssd_gpu.so (the product) and libssd_scene.so (synthetic input source for tests / bench)."""
    _build_one(SCENE_LIB, SCENE_SOURCES, force, verbose)
    return _build_one(LIB, SOURCES, force, verbose, log="ptxas.log")


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
