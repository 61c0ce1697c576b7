"""``MeanFlagger`` -- flag channels whose sigma exceeds mean + nsigma * standard error
(reference: transformers/flaggers/mean_flagger.py:10-46)."""
from dataclasses import dataclass

import numpy as np

from ...base.dataset import Dataset
from .flagger import Flagger


def normal_flagging(x, mean_sigma=None, nsigma=0.0):
    """(preserved indices, outlier indices) of a 1-D array; threshold = mean + nsigma * std / sqrt(n)"""
    x = np.asarray(x)
    center = np.mean(x) if mean_sigma is None else mean_sigma
    threshold = center + nsigma * np.std(x) / np.sqrt(len(x))
    return np.where(x <= threshold)[0], np.where(x > threshold)[0]


@dataclass(init=True, repr=True)
class MeanFlagger(Flagger):

    def run(self, mean_sigma=None, nsigma=0.0):
        if self.nsigma is not None:
            nsigma = self.nsigma
        if not isinstance(self.dataset, Dataset):
            raise TypeError("The data attribute is not a Dataset")
        sigma = np.array(self.dataset.sigma, dtype=np.float64, copy=True)
        if sigma.ndim == 1:
            keep, outliers = normal_flagging(sigma, mean_sigma, nsigma)
            self._apply(keep, outliers, sigma)
            return keep, outliers
        mask = self._per_pixel_mask(sigma, mean_sigma, nsigma)
        self._apply(None, mask, sigma)
        return ~mask, mask

    def _per_pixel_mask(self, sigma, mean_sigma, nsigma):
        """outlier mask of a per-pixel sigma ``(m, *pix)``: on the GPU when there is one (csr_flag_mean), NumPy otherwise"""
        import torch
        if torch.cuda.is_available():
            from .flagger import device_flag
            return device_flag(sigma, np.ones(sigma.shape[0]), "mean", nsigma, mean_sigma=mean_sigma)[1]
        center = sigma.mean(axis=0) if mean_sigma is None else np.asarray(mean_sigma)
        return sigma > center + nsigma * sigma.std(axis=0) / np.sqrt(sigma.shape[0])
