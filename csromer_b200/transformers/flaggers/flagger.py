"""``Flagger`` -- base of the channel flaggers (reference: transformers/flaggers/flagger.py:9-41).

A flagger looks at ``dataset.sigma`` and either zero-weights the outlier channels in place (``dataset.w[idx] = 0``,
which -- like the reference -- leaves ``k`` / ``theo_noise`` untouched) or deletes them (``delete_channels=True``:
``lambda2`` / ``sigma`` / ``data`` are re-assigned, so the dataset gets a new sampling).  The weights they produce are
what the hot path consumes as per-channel / per-pixel flags; a stack of datasets whose channels were deleted
differently goes to ``PerPixelDFT1D``.

Batch extension: ``sigma`` may be ``(m, *pix)``; the statistic is then taken per pixel along the channel axis and only
the zero-weighting mode is available (deleting would give every pixel its own channel count).
"""
from abc import ABCMeta, abstractmethod
from dataclasses import dataclass

import numpy as np

from ...base.dataset import Dataset


def median_absolute_deviation(x):
    """median |x - median(x)| (flagger.py:11-17)"""
    x = np.asarray(x)
    return np.median(np.abs(x - np.median(x)))


def moving_average(x, w):
    """box average over ``w`` samples, ``valid`` part only (flagger.py:20-21)"""
    return np.convolve(x, np.ones(int(w)), "valid") / w


@dataclass(init=True, repr=True)
class Flagger(metaclass=ABCMeta):
    dataset: Dataset = None
    nsigma: float = None
    delete_channels: bool = None

    def __post_init__(self):
        if self.delete_channels is None:
            self.delete_channels = False

    @abstractmethod
    def run(self):
        pass

    # shared tail of every run(): apply the decision to the dataset
    def _apply(self, keep, outliers, sigma):
        ds = self.dataset
        total = sigma.shape[0] if sigma.ndim == 1 else sigma.size
        flagged = np.count_nonzero(outliers) if outliers.dtype == bool else len(outliers)
        if sigma.ndim > 1:
            if self.delete_channels:
                raise ValueError("delete_channels needs a single line of sight; zero-weight the channels instead or "
                                 "build one Dataset per pixel and use PerPixelDFT1D")
            if np.ndim(ds.w) != sigma.ndim:
                ds.w = np.broadcast_to(np.reshape(ds.w, (-1,) + (1,) * (sigma.ndim - 1)), sigma.shape).copy()
            ds.w[outliers] = 0.0
        elif self.delete_channels:
            data = ds.data
            ds.lambda2 = ds.lambda2[keep]
            ds.sigma = sigma[keep]
            if data is not None:
                ds.data = data[keep]
        else:
            ds.w[outliers] = 0.0
        print("Flagging {0:.2f}% of the data".format(100.0 * flagged / total))


def device_flag(sigma, weights, kind, nsigma, mean_sigma=None, window=None):
    """Per-pixel flagging on the GPU (csrc/flag.cu).  ``sigma`` ``(m, *pix)`` host array or CUDA tensor, ``weights`` like it
    (or ``(m,)``, broadcast); returns (new weights as a float32 array / tensor of sigma's shape, boolean outlier mask of the
    same type).  ``kind``: "mean" (MeanFlagger) or "hampel" (HampelFlagger, ``window`` channels)."""
    import torch

    from ... import _device as dv
    from ... import _lib
    lib = _lib.load()
    on_device = dv.is_torch(sigma) and sigma.is_cuda
    m = sigma.shape[0]
    pix = tuple(sigma.shape[1:])
    sig, _ = dv.pack_real(sigma, m, m)                                           # [P, m]
    P = sig.shape[0]
    marks = torch.ones_like(sig)                                                 # the kernel zeroes the outlier channels
    if kind == "mean":
        center = None if mean_sigma is None else dv.per_pixel_vector(mean_sigma, P, sig.device, "mean_sigma")
        _lib.check(lib.csr_flag_mean(dv.tptr(sig), P, m, dv.tptr(center), float(nsigma), dv.tptr(marks), None, dv.stream_ptr()),
                   "csr_flag_mean")
    elif kind == "hampel":
        _lib.check(lib.csr_flag_hampel(dv.tptr(sig), P, m, int(window), float(nsigma), dv.tptr(marks), None, dv.stream_ptr()),
                   "csr_flag_hampel")
    else:
        raise ValueError("kind must be 'mean' or 'hampel'")
    out_mask = dv.unpack_real(marks, m, pix, like_torch=True) == 0
    w_t = dv.to_device(weights).to(torch.float32)
    if w_t.dim() == 1:
        w_t = w_t.reshape((m,) + (1,) * len(pix)).expand((m,) + pix)
    new_w = torch.where(out_mask, torch.zeros((), device=w_t.device), w_t.reshape((m,) + pix))
    if on_device:
        return new_w, out_mask
    return new_w.cpu().numpy(), out_mask.cpu().numpy()
