"""Action normalizers — mirror of judo/utils/normalization.py:78-222 (host-side; default is "none",
judo/controller/controller.py:42).  Applied around sampling in Controller.update_action."""

from __future__ import annotations

import warnings
from abc import ABC, abstractmethod

import numpy as np


class Normalizer(ABC):
    def __init__(self, dim: int) -> None:
        self.dim = dim

    @abstractmethod
    def normalize(self, x: np.ndarray) -> np.ndarray: ...

    @abstractmethod
    def denormalize(self, x: np.ndarray) -> np.ndarray: ...

    def update(self, x: np.ndarray) -> None:
        return None


class IdentityNormalizer(Normalizer):
    def normalize(self, x: np.ndarray) -> np.ndarray:
        return x

    def denormalize(self, x: np.ndarray) -> np.ndarray:
        return x


class MinMaxNormalizer(Normalizer):
    """Affine map of [min, max] onto [-1, 1]; dimensions with an infinite bound pass through."""

    def __init__(self, dim: int, min: np.ndarray, max: np.ndarray, eps: float = 1e-6) -> None:  # noqa: A002
        super().__init__(dim)
        self.min, self.max, self.eps = min, max, eps
        finite = (self.min != -np.inf) & (self.max != np.inf)
        self.norm_dims = np.where(finite)[0]
        if len(self.norm_dims) != dim:
            warnings.warn(f"MinMaxNormalizer: action dimensions {np.where(~finite)[0].tolist()} have infinite range and will "
                          "not be normalized.", UserWarning, stacklevel=2)

    def normalize(self, x: np.ndarray) -> np.ndarray:
        out = x.copy()
        lo, hi = self.min[self.norm_dims], self.max[self.norm_dims]
        out[..., self.norm_dims] = 2 * (x[..., self.norm_dims] - lo) / (hi - lo) - 1
        return out

    def denormalize(self, x: np.ndarray) -> np.ndarray:
        out = x.copy()
        lo, hi = self.min[self.norm_dims], self.max[self.norm_dims]
        out[..., self.norm_dims] = (x[..., self.norm_dims] + 1) * (hi - lo) / 2 + lo
        return out


class RunningMeanStdNormalizer(Normalizer):
    """Per-dimension running mean / std with the batched Welford update of normalization.py:170-193."""

    def __init__(self, dim: int, init_std: float = 1.0, min_std: float = 1e-5, max_std: float = 1e3, eps: float = 1e-6) -> None:
        super().__init__(dim)
        self.eps, self.min_std, self.max_std = eps, min_std, max_std
        self.count = 0
        self.mean = np.zeros(dim)
        self.std = np.ones(dim) * init_std
        self.M2 = np.zeros(dim)

    def update(self, x: np.ndarray) -> None:
        assert x.shape[-1] == self.dim, f"Expected dimension {self.dim}, but got {x.shape[-1]}"
        axes = tuple(range(x.ndim - 1))
        self.count += np.prod(x.shape[:-1])
        delta = x - self.mean
        self.mean += np.sum(delta, axis=axes) / self.count
        self.M2 = np.maximum(self.M2 + np.sum(delta * (x - self.mean), axis=axes), 0)
        self.std = np.clip(np.sqrt(self.M2 / self.count), self.min_std, self.max_std)

    def normalize(self, x: np.ndarray) -> np.ndarray:
        return (x - self.mean) / (self.std + self.eps)

    def denormalize(self, x: np.ndarray) -> np.ndarray:
        return x * self.std + self.mean


normalizer_registry = {"none": IdentityNormalizer, "min_max": MinMaxNormalizer, "running": RunningMeanStdNormalizer}


def make_normalizer(normalizer_type: str, dim: int, **kwargs) -> Normalizer:  # noqa: ANN003
    if normalizer_type not in normalizer_registry:
        raise ValueError(f"Invalid normalizer type: {normalizer_type}")
    return normalizer_registry[normalizer_type](dim, **kwargs)
