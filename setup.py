"""In-tree build of the native module: ``python setup.py build_ext --inplace``.

Produces ``shallowspeed_b200/_C*.so`` (sm_100a only).  The .so is git-ignored but travels
with the gpurun snapshot, so the GPU box never has to compile."""
import glob
import os

from setuptools import find_packages, setup
from torch.utils.cpp_extension import BuildExtension, CUDAExtension

ROOT = os.path.dirname(os.path.abspath(__file__))
NCCL_ROOT = None
try:
    import nvidia.nccl as _nccl

    NCCL_ROOT = os.path.dirname(_nccl.__file__)
except Exception:
    pass

sources = (["csrc/bindings.cpp"] + sorted(glob.glob("csrc/kernels/*.cu")) + sorted(glob.glob("csrc/runtime/*.cpp"))
           + sorted(glob.glob("csrc/runtime/*.cu")))
include_dirs = [os.path.join(ROOT, "csrc")]
library_dirs, libraries, extra_link = [], [], []
if NCCL_ROOT:
    include_dirs.append(os.path.join(NCCL_ROOT, "include"))
    extra_link += [f"-L{os.path.join(NCCL_ROOT, 'lib')}", "-l:libnccl.so.2", f"-Wl,-rpath,{os.path.join(NCCL_ROOT, 'lib')}"]

os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 4))

setup(
    name="shallowspeed_b200",
    version="0.1.0",
    packages=find_packages(include=["shallowspeed_b200*"]),
    ext_modules=[
        CUDAExtension(
            name="shallowspeed_b200._C",
            sources=sources,
            include_dirs=include_dirs,
            extra_compile_args={
                "cxx": ["-O3", "-std=c++17"],
                "nvcc": ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
                         "--expt-relaxed-constexpr"],
            },
            extra_link_args=extra_link,
        )
    ],
    cmdclass={"build_ext": BuildExtension.with_options(use_ninja=True)},
)
