"""mslesions3d_b200 -- the SSD3D detector hot path of MSLesions3D on hand-written sm_100a CUDA.

Layout:
  csrc/            CUDA kernels + the C ABI (include/ssd3d_b200.h), built into libssd3d_b200.so
  _lib, ops        ctypes binding and tensor-level launch wrappers
  mobilenet, ssd3d, utils, predict   the reference's module/class/function surface on top of the kernels
  training         train-mode forward with a tape, hand-written backward, flat-buffer Adam, captured fit_step
  parallel         batch-of-volumes data parallelism helpers (torchrun / NCCL, gloo in CPU tests)
  synthetic        in-memory synthetic lesion volumes and random-init weights (the benchmark input spec)
"""
__version__ = "0.1.0"


def __getattr__(name):  # lazy: importing the package must not need torch-cuda or the built library
    if name in ("LSSD3D", "SSD3D", "MobileNetBase", "PredictionConvolutions", "MultiBoxLoss"):
        from . import ssd3d
        return getattr(ssd3d, name)
    raise AttributeError(name)
