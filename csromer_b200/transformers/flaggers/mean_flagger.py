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
        center = sigma.mean(axis=0) if mean_sigma is None else np.asarray(mean_sigma)
        mask = sigma > center + nsigma * sigma.std(axis=0) / np.sqrt(sigma.shape[0])
        self._apply(None, mask, sigma)
        return ~mask, mask
