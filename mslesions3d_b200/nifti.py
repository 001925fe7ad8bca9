"""Minimal NIfTI-1 single-file (``.nii`` / ``.nii.gz``) writer and reader.

The reference stores its prediction volumes with ``nib.Nifti1Image(array, affine)`` + ``nib.save``
(predict.py:225-226) and its synthetic data set the same way (generate_artificial_dataset.py:96-103);
nibabel is not installed here, so the on-disk format is written directly: the 348-byte NIfTI-1 header
(``sizeof_hdr`` 348, magic ``n+1``, ``vox_offset`` 352), four zero extension bytes, then the voxels in
Fortran order (first array axis fastest), as nibabel does for a C-ordered numpy array.  The affine goes
into the sform (``sform_code`` 2, "aligned"), which is what nibabel writes for an affine passed to the
constructor; ``pixdim`` holds the column norms of the affine.  Only what this package needs: 3-D / 4-D
arrays of uint8, int16, int32, float32, float64, uint16.
"""
from __future__ import annotations

import gzip
import struct

import numpy as np

# NIfTI-1 datatype codes (nifti1.h)
_DTYPE_CODES = {np.dtype("uint8"): 2, np.dtype("int16"): 4, np.dtype("int32"): 8, np.dtype("float32"): 16,
                np.dtype("float64"): 64, np.dtype("uint16"): 512}
_CODE_DTYPES = {v: k for k, v in _DTYPE_CODES.items()}


def _header(shape, dtype, affine) -> bytes:
    ndim = len(shape)
    if not 1 <= ndim <= 7:
        raise ValueError("NIfTI-1 holds 1 to 7 dimensions, got %d" % ndim)
    dt = np.dtype(dtype)
    if dt not in _DTYPE_CODES:
        raise ValueError("unsupported dtype %s" % dt)
    affine = np.asarray(affine, dtype=np.float64).reshape(4, 4)
    dim = [ndim] + list(shape) + [1] * (7 - ndim)
    zooms = np.sqrt((affine[:3, :3] ** 2).sum(0))
    pixdim = [1.0] + [float(z) for z in zooms[:min(3, ndim)]] + [1.0] * (7 - min(3, ndim))
    h = bytearray(348)
    struct.pack_into("<i", h, 0, 348)                       # sizeof_hdr
    struct.pack_into("<8h", h, 40, *dim)                    # dim[8]
    struct.pack_into("<h", h, 70, _DTYPE_CODES[dt])         # datatype
    struct.pack_into("<h", h, 72, dt.itemsize * 8)          # bitpix
    struct.pack_into("<8f", h, 76, *pixdim)                 # pixdim[8] (qfac = 1)
    struct.pack_into("<f", h, 108, 352.0)                   # vox_offset
    struct.pack_into("<f", h, 112, 1.0)                     # scl_slope
    struct.pack_into("<f", h, 116, 0.0)                     # scl_inter
    h[123] = 2                                              # xyzt_units: millimetres
    struct.pack_into("<h", h, 252, 0)                       # qform_code
    struct.pack_into("<h", h, 254, 2)                       # sform_code: aligned
    struct.pack_into("<4f", h, 280, *affine[0])             # srow_x
    struct.pack_into("<4f", h, 296, *affine[1])             # srow_y
    struct.pack_into("<4f", h, 312, *affine[2])             # srow_z
    h[344:348] = b"n+1\x00"                                 # magic: header and data in one file
    return bytes(h)


def save_nifti(path: str, array, affine=None) -> None:
    """Write ``array`` (numpy, C order, axes x, y, z[, t]) with ``affine`` (4x4, identity if None)."""
    a = np.asarray(array)
    if affine is None:
        affine = np.eye(4)
    blob = _header(a.shape, a.dtype, affine) + b"\x00" * 4 + np.asfortranarray(a).tobytes(order="F")
    if str(path).endswith(".gz"):
        with gzip.open(path, "wb", compresslevel=1) as f:
            f.write(blob)
    else:
        with open(path, "wb") as f:
            f.write(blob)


def load_nifti(path: str):
    """-> (array in C order with axes x, y, z[, t], 4x4 float64 affine from the sform)."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        blob = f.read()
    if struct.unpack_from("<i", blob, 0)[0] != 348 or blob[344:347] != b"n+1":
        raise ValueError("%s is not a little-endian single-file NIfTI-1 volume" % path)
    dim = struct.unpack_from("<8h", blob, 40)
    shape = tuple(dim[1:1 + dim[0]])
    code = struct.unpack_from("<h", blob, 70)[0]
    if code not in _CODE_DTYPES:
        raise ValueError("unsupported NIfTI datatype code %d" % code)
    off = int(struct.unpack_from("<f", blob, 108)[0])
    dt = _CODE_DTYPES[code]
    n = int(np.prod(shape))
    data = np.frombuffer(blob, dtype=dt, count=n, offset=off).reshape(shape, order="F")
    slope, inter = struct.unpack_from("<2f", blob, 112)
    if slope not in (0.0, 1.0) or inter != 0.0:
        data = data * slope + inter
    affine = np.eye(4)
    for r, o in enumerate((280, 296, 312)):
        affine[r] = struct.unpack_from("<4f", blob, o)
    return np.ascontiguousarray(data), affine
