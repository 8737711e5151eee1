"""Profile -> autoencoder hand-off on the device.

The reference writes com_profs / cov_profs as text, parses them back with float() (pipelines.py:310-330), saves
.npy, reloads them in ae_utils.vae_encode (ae_utils.py:299-301) and min-max scales + casts to float32 in
make_data_loader (ae_utils.py:19-32).  Here the integer profiles that are already in HBM become the very same
float32 tensors without leaving the GPU.  PyTorch appears only at this seam, where tensors are handed to the
existing autoencoder; the values themselves come from the CUDA kernel behind lrb_dev_profile_values.
"""
import ctypes as C

from ._lib import check, lib
from .profile import COMP_WIDTH, _stream


def profile_values(counts, denom, k=0):
    """float64 [N, width] == float(the '%f' text) of every profile value.  counts: int32 [N, width] CUDA tensor;
    denom: int32 [N] (read lengths for composition with k = 3/4/5, window sums for coverage with k = 0)."""
    import torch
    assert counts.is_cuda and counts.dtype == torch.int32 and counts.is_contiguous() and denom.is_contiguous()
    n, width = counts.shape
    if k:
        assert COMP_WIDTH[k] == width
    out = torch.empty((n, width), dtype=torch.float64, device=counts.device)
    check(lib.lrb_dev_profile_values(C.c_void_p(counts.data_ptr()), C.c_void_p(denom.data_ptr()), n, width, int(k),
                                     C.c_void_p(out.data_ptr()), _stream()))
    return out


def minmax_scale(x):
    """sklearn.preprocessing.MinMaxScaler().fit_transform(x), feature_range (0, 1), in float64, operation for operation
    (scale = 1 / range with ranges below 10 eps replaced by 1; x * scale + (0 - min * scale))."""
    import torch
    mn, mx = x.min(dim=0).values, x.max(dim=0).values
    rng = mx - mn
    rng = torch.where(rng < 10 * torch.finfo(torch.float64).eps, torch.ones_like(rng), rng)
    scale = 1.0 / rng
    return x * scale + (0.0 - mn * scale)


def vae_inputs(comp_counts, read_len, hist, sums, k):
    """(covs, profs) float32 CUDA tensors — what make_data_loader (ae_utils.py:19-32) builds from the .npy files."""
    profs = minmax_scale(profile_values(comp_counts, read_len, k)).float()
    covs = minmax_scale(profile_values(hist, sums, 0)).float()
    return covs, profs
