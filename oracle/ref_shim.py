"""Import the UNMODIFIED reference (``lesions3d/{ssd3d,mobilenet,utils,base_network}.py``) on CPU.

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference depends on pytorch_lightning / monai /
matplotlib, none of which are installed; this shim registers inert stand-ins
for the names the reference touches at import time and
then imports ``ssd3d``, ``mobilenet`` and ``utils``.

Where the modules come from: the read-only mount ``/root/reference/lesions3d`` in the build
container, else the byte-identical copies ``oracle/make_ref.py`` staged under ``oracle/_ref/``
(git-ignored; they travel to the GPU box, where the mount does not exist).  Users:
``tests/golden/make_golden.py`` (produces the committed golden vectors),
``tests/test_oracle_vs_reference.py`` (live oracle-vs-reference comparison) and the
``--impl reference`` arm of ``bench.py`` (times the reference's own forward + detect_objects).
The GPU parity tests and ``smoke()`` never touch it.
"""
import importlib
import os
import sys
import types

_MOUNT = "/root/reference/lesions3d"
_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "lesions3d")
REFERENCE_DIR = os.environ.get("MSL3D_REFERENCE_DIR") or (
    _MOUNT if os.path.isfile(os.path.join(_MOUNT, "ssd3d.py")) else _STAGED)


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "ssd3d.py"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    import torch.nn as nn

    class _LightningModule(nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

        @property
        def device(self):
            import torch
            return torch.device("cpu")

    _mod("pytorch_lightning", LightningModule=_LightningModule, LightningDataModule=object)

    class _A:  # two distinct dummy bases (utils.py:398 inherits from both)
        def __init__(self, *a, **k):
            pass

    class _B:
        pass

    monai = _mod("monai")
    monai.losses = _mod("monai.losses", FocalLoss=object)
    monai.config = _mod("monai.config", KeysCollection=object)
    _mod("monai.config.type_definitions", NdarrayOrTensor=object)
    monai.transforms = _mod("monai.transforms")
    _mod("monai.transforms.transform", MapTransform=_A)
    _mod("monai.transforms.inverse", InvertibleTransform=_B)
    def _box_area(boxes):      # monai.data.box_area: product of the (max - min) extents of [mins..., maxs...] boxes
        nd = boxes.shape[1] // 2
        area = boxes[:, nd] - boxes[:, 0]
        for ax in range(1, nd):
            area = area * (boxes[:, ax + nd] - boxes[:, ax])
        return area

    monai.data = _mod("monai.data", box_area=_box_area)
    monai.networks = _mod("monai.networks")
    _mod("monai.networks.blocks", Convolution=object)
    mpl = _mod("matplotlib")
    mpl.pyplot = _mod("matplotlib.pyplot")
    _mod("mpl_toolkits")
    _mod("mpl_toolkits.axes_grid1", make_axes_locatable=lambda *a, **k: None)
    _mod("wandb")


_cached = None


def load_reference():
    """Return the (ssd3d, mobilenet, utils) modules of the unmodified reference."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise RuntimeError("reference sources not present at %s" % REFERENCE_DIR)
    saved = {k: sys.modules.get(k) for k in ("utils", "mobilenet", "ssd3d", "base_network")}
    _install_stubs()
    sys.path.insert(0, REFERENCE_DIR)
    try:
        for k in saved:
            sys.modules.pop(k, None)
        ssd3d = importlib.import_module("ssd3d")
        mobilenet = importlib.import_module("mobilenet")
        utils = importlib.import_module("utils")
    finally:
        sys.path.remove(REFERENCE_DIR)
    _cached = (ssd3d, mobilenet, utils)
    return _cached
