"""Loader-side fusion: the one arithmetic step of das4whales.data_handle that sits between the
file and the hot path.  File parsing (HDF5 / TDMS) stays with the reference's loaders."""
import numpy as np

from . import rows as _rows
from .dsp import _is_tensor, _to_host64


def raw2strain(trace, metadata):
    """(trace - mean over time of each channel) * metadata["scale_factor"]  (reference: data_handle.py:157-177).

    ndarray in -> float64 ndarray out (float arrays are updated in place like the reference, which cannot take integer
    arrays at all); CUDA tensor (int32 raw counts as stored on disk, or float32) in -> float32 CUDA tensor out, so a
    file can be uploaded as int32 and converted on the device."""
    scale = float(metadata["scale_factor"])
    if _is_tensor(trace):
        return _rows.raw2strain(trace, scale)
    import torch
    arr = np.asarray(trace)
    if arr.dtype == np.int32:
        dev = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
    else:
        dev = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).cuda()
    out = _to_host64(_rows.raw2strain(dev, scale))
    if isinstance(trace, np.ndarray) and np.issubdtype(trace.dtype, np.floating):
        trace[...] = out                     # the reference works in place (trace -= mean; trace *= scale_factor)
        return trace
    return out
