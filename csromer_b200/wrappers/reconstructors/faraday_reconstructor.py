"""``FaradayReconstructorWrapper`` -- abstract base of the reconstruction wrappers
(reference: wrappers/reconstructors/faraday_reconstructor.py:9-23)."""
from abc import ABCMeta, abstractmethod
from dataclasses import dataclass

from ...base.dataset import Dataset


@dataclass(init=True, repr=True)
class FaradayReconstructorWrapper(metaclass=ABCMeta):
    dataset: Dataset = None

    @abstractmethod
    def config_fd_space(self):
        pass

    @abstractmethod
    def reconstruct(self):
        pass

    @abstractmethod
    def calculate_second_moment(self):
        pass
