"""``HampelFlagger`` -- robust outlier flagging of sigma against the median of its moving average
(reference: transformers/flaggers/hampel_flagger.py:9-60).

The reference's ``run`` passes ``dataset.w`` (the weight ARRAY) where ``hampel`` expects the window length
(hampel_flagger.py:40,48) and cannot work as written; the window ``w`` given to the
constructor is used here.
"""
from dataclasses import dataclass

import numpy as np

from ...base.dataset import Dataset
from .flagger import Flagger, median_absolute_deviation, moving_average


def hampel(array, w, nsigma, imputation=False):
    """indices within / beyond nsigma * 1.4826 * MAD of the moving average's median; optionally impute the outliers"""
    values = np.array(array, dtype=np.float64, copy=True)
    smooth = moving_average(values, w)
    center = np.median(smooth)
    spread = 1.4826 * median_absolute_deviation(smooth)
    dist = np.abs(values - center)
    keep, outliers = np.where(dist <= nsigma * spread)[0], np.where(dist > nsigma * spread)[0]
    if imputation:
        values[outliers] = center
        return values, keep, outliers
    return keep, outliers


@dataclass(init=True, repr=True)
class HampelFlagger(Flagger):
    w: int = None
    imputation: bool = None

    def run(self, nsigma=0.0):
        if self.nsigma is not None:
            nsigma = self.nsigma
        if not isinstance(self.dataset, Dataset):
            raise TypeError("The data attribute is not a Dataset")
        if self.w is None:
            raise ValueError("HampelFlagger needs the window length w")
        sigma = np.array(self.dataset.sigma, dtype=np.float64, copy=True)
        if sigma.ndim != 1:  # batch extension: every pixel its own decision, weights zeroed (csr_flag_hampel / NumPy)
            if self.imputation:
                raise ValueError("imputation works on a single line of sight (shared sigma)")
            import torch
            if torch.cuda.is_available():
                from .flagger import device_flag
                mask = device_flag(sigma, np.ones(sigma.shape[0]), "hampel", nsigma, window=self.w)[1]
            else:
                flat = sigma.reshape(sigma.shape[0], -1)
                mask = np.zeros(flat.shape, dtype=bool)
                for p in range(flat.shape[1]):
                    mask[hampel(flat[:, p], self.w, nsigma, False)[1], p] = True
                mask = mask.reshape(sigma.shape)
            self._apply(None, mask, sigma)
            return ~mask, mask
        if self.imputation:
            new_sigma, keep, outliers = hampel(sigma, self.w, nsigma, True)
            self.dataset.sigma = new_sigma
            print("Imputing {0:.2f}% of the data".format(100.0 * len(outliers) / len(sigma)))
            return None
        keep, outliers = hampel(sigma, self.w, nsigma, False)
        self._apply(keep, outliers, sigma)
        return keep, outliers
