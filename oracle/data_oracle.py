"""TEST INFRASTRUCTURE ONLY -- CPU restatement (float64 NumPy) of the one arithmetic step of the reference's loader
module that the GPU path fuses into the upload: data_handle.raw2strain (data_handle.py:157-177).  Pinned against the
unmodified reference in oracle/make_golden.py (tests/golden/raw2strain.npz).  Only tests/, __graft_entry__.smoke()
and bench.py's CPU legs may import this."""
import numpy as np


def raw2strain(trace, metadata):
    """trace -= mean(trace, axis=1, keepdims=True); trace *= metadata["scale_factor"]   (data_handle.py:175-176).
    The reference updates a float array in place and returns it; this restatement returns a new float64 array (and
    accepts the on-disk int32 counts, which the in-place reference cannot)."""
    t = np.array(trace, dtype=np.float64)
    t -= np.mean(t, axis=1, keepdims=True)
    t *= metadata["scale_factor"]
    return t
