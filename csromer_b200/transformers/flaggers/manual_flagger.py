"""``ManualFlagger`` -- flag the channels the caller lists (reference: transformers/flaggers/manual_flagger.py:12-47)."""
import numpy as np

from ...base.dataset import Dataset
from .flagger import Flagger


class ManualFlagger(Flagger):

    def __init__(self, dataset=None, nsigma=None, delete_channels=None, outlier_idxs=None):
        super().__init__(dataset=dataset, nsigma=nsigma, delete_channels=delete_channels)
        self.outlier_idxs = [] if outlier_idxs is None else outlier_idxs

    def run(self):
        if not isinstance(self.dataset, Dataset):
            raise TypeError("The data attribute is not a Dataset")
        sigma = np.array(self.dataset.sigma, dtype=np.float64, copy=True)
        if sigma.ndim != 1:
            raise ValueError("ManualFlagger works on a shared sigma")
        outliers = np.atleast_1d(np.asarray(self.outlier_idxs, dtype=np.int64))
        keep = np.setxor1d(np.arange(len(sigma)), outliers)
        self._apply(keep, outliers, sigma)
        return keep, self.outlier_idxs
