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
