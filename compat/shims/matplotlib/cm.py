class ScalarMappable:
    def __init__(self, norm=None, cmap=None):
        self.norm, self.cmap = norm, cmap

    def to_rgba(self, x):
        return self.cmap(self.norm(x) if self.norm is not None else x)
