import numpy as np


class _Ramp:
    """a two-colour ramp standing in for a named colormap (only used for TensorBoard images)."""

    def __init__(self, name):
        self.name = name

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float64)
        if self.name.endswith("_r"):
            x = 1.0 - x
        return np.stack([x, 1.0 - np.abs(2.0 * x - 1.0), 1.0 - x, np.ones_like(x)], axis=-1)


def get_cmap(name):
    return _Ramp(name)
