"""Drop-in for the reference's compiled _track module (/root/reference/_track.pyx:13-25):
runSum on the GPU.  Host arrays in and out, like the Cython function."""
import numpy as np

from . import tracks_device


def runSum(mask, outSum):
    """outSum[i] = number of zeros in mask[0:i] (exclusive running count; _track.pyx:13-25)"""
    assert mask.dtype == np.uint8 and outSum.dtype == np.int32, "Buffer dtype mismatch"
    assert len(mask) == len(outSum)
    if len(mask) == 0:
        return
    import torch
    d = torch.from_numpy(np.ascontiguousarray(mask)).to(torch.device("cuda", tracks_device._ctx().device))
    outSum[:] = tracks_device.run_sum(d).cpu().numpy()
