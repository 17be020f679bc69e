class ImageDataset:
    def __init__(self, *a, **k):
        raise NotImplementedError("timm.data.dataset.ImageDataset is imported but never used by the TULIP driver")
