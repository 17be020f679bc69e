"""Pin oracle/input_pipeline.py against the UNMODIFIED transform classes of tulip/util/datasets.py, composed as the three dataset
builders compose them, and write tests/golden/input_pipeline.npz.   python -m oracle.make_golden_input   (build container only)"""
import os
import sys
import types

import numpy as np
import torch

from . import input_pipeline as P

REF_ROOT = os.environ.get("TULIP_REFERENCE", "/root/reference")
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def import_reference_datasets():
    if "timm" not in sys.modules or not hasattr(sys.modules["timm"], "data"):          # timm is not installed; unused by the transforms
        timm = sys.modules.get("timm") or types.ModuleType("timm")
        data, const, ds = (types.ModuleType(n) for n in ("timm.data", "timm.data.constants", "timm.data.dataset"))
        data.create_transform = None
        const.IMAGENET_DEFAULT_MEAN = const.IMAGENET_DEFAULT_STD = None
        ds.ImageDataset = object
        timm.data, data.constants, data.dataset = data, const, ds
        sys.modules.update({"timm": timm, "timm.data": data, "timm.data.constants": const, "timm.data.dataset": ds})
    sys.path.insert(0, os.path.join(REF_ROOT, "tulip"))
    import util.datasets as D
    return D


def reference_chain(D, frame, dataset, out_size, in_size, log_transform):
    from torchvision import transforms
    scale, fmin = {"kitti": (1 / 80, None), "durlar": (1 / 120, 0.3 / 120), "carla": (1 / 80, 2 / 80)}[dataset]
    t_low, t_high = [transforms.ToTensor(), D.ScaleTensor(scale)], [transforms.ToTensor(), D.ScaleTensor(scale)]
    if fmin is not None:
        t_low.append(D.FilterInvalidPixels(min_range=fmin, max_range=1)); t_high.append(D.FilterInvalidPixels(min_range=fmin, max_range=1))
    t_low.append(D.DownsampleTensor(h_high_res=out_size[0], downsample_factor=out_size[0] // in_size[0]))
    if out_size[1] // in_size[1] > 1:
        t_low.append(D.DownsampleTensorWidth(w_high_res=out_size[1], downsample_factor=out_size[1] // in_size[1]))
    if log_transform:
        t_low.append(D.LogTransform()); t_high.append(D.LogTransform())
    rng_map = frame[..., 0].astype(np.float32)                   # npy_loader
    return transforms.Compose(t_low)(rng_map), transforms.Compose(t_high)(rng_map)


def main():
    D = import_reference_datasets()
    rng = np.random.Generator(np.random.PCG64(77))
    out = {}
    for dataset, out_size, in_size in (("kitti", (64, 1024), (16, 1024)), ("durlar", (128, 2048), (32, 2048)), ("carla", (64, 1024), (16, 512))):
        raw = (rng.random((2, *out_size, 2), dtype=np.float32) * 130.0).astype(np.float32)      # metres, some beyond the sensor range
        raw[rng.random(raw.shape) < 0.03] = 0.0
        for log_t in (True, False):
            lo_o, hi_o = P.preprocess(raw, dataset, in_size[0], in_size[1], log_t)
            for b in range(2):
                lo_r, hi_r = reference_chain(D, raw[b], dataset, out_size, in_size, log_t)
                assert torch.equal(lo_r, lo_o[b]) and torch.equal(hi_r, hi_o[b]), (dataset, log_t)
        out[f"{dataset}_raw"] = raw[:, ::8, ::16].copy()
        lo_o, hi_o = P.preprocess(raw, dataset, in_size[0], in_size[1], True)
        out[f"{dataset}_hi"] = hi_o.numpy()[:, :, ::8, ::16].copy()
        out[f"{dataset}_lo_sum"] = np.array([float(lo_o.double().sum())])
    # the CARLA .rimg container: the reference's own loader on files written here (datasets.py:181-193)
    import tempfile
    for name, (s0, s1) in (("rimg_a", (16, 48)), ("rimg_b", (64, 1024))):
        payload = (rng.random((s1, s0), dtype=np.float32) * 90.0).astype(np.float16)
        payload[rng.random(payload.shape) < 0.05] = 0
        buf = np.array([s0, s1], dtype=np.uint64).tobytes() + payload.tobytes()
        with tempfile.NamedTemporaryFile(suffix=".rimg") as f:
            f.write(buf); f.flush()
            ref = D.rimg_loader(f.name)
        got, stored = P.rimg_decode(buf)
        assert ref.dtype == np.float32 and ref.shape == (s0, s1) and np.array_equal(ref, got), name
        assert np.array_equal(stored, payload)
        if name == "rimg_a":
            out["rimg_a_file"] = np.frombuffer(buf, dtype=np.uint8).copy()
            out["rimg_a_frame"] = ref.copy()
        else:
            out["rimg_b_seed"] = np.array([s0, s1], dtype=np.int64)
            out["rimg_b_frame_sum"] = np.array([float(ref.astype(np.float64).sum()), float(ref[3, 17]), float(ref[-1, 0])])
            lo_o, hi_o = P.preprocess(ref[None, :, :, None], "carla", 16, 1024, True)
            out["rimg_b_hi_sum"] = np.array([float(hi_o.double().sum()), float(lo_o.double().sum())])
            out["rimg_b_file"] = np.frombuffer(buf, dtype=np.uint8).copy()
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, "input_pipeline.npz"), **out)
    print("input pipeline oracle pinned bit for bit against the reference transform chains (kitti, durlar, carla; log and linear) "
          "and rimg_loader")


if __name__ == "__main__":
    main()
