import numpy as np


class Normalize:
    def __init__(self, vmin=0.0, vmax=1.0):
        self.vmin, self.vmax = vmin, vmax

    def __call__(self, x):
        return np.clip((np.asarray(x, dtype=np.float64) - self.vmin) / (self.vmax - self.vmin), 0.0, 1.0)
