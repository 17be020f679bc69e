"""The UNMODIFIED reference driver (baseline/_ref/tulip/main_lidar_upsampling.py -> engine_upsampling.py, staged by
__graft_entry__.build()) against the drop-in module.  SURVEY 4 (vi) / 8b: `import model.tulip as tulip` must resolve to
tulip_b200 when compat/dropin is ahead on sys.path, and the driver must train, checkpoint, resume and evaluate unchanged."""
import glob
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from compat import env as cenv

needs_ref = pytest.mark.skipif(not cenv.reference_available(), reason="baseline/_ref not staged (run __graft_entry__.build() where /root/reference exists)")

ARGS = ["--model_select", "tulip_base", "--window_size", "2", "8", "--patch_size", "1", "4", "--pixel_shuffle", "--circular_padding",
        "--patch_unmerging", "--log_transform", "--dataset_select", "kitti", "--img_size_low_res", "16", "1024",
        "--img_size_high_res", "64", "1024", "--in_chans", "1", "--batch_size", "2", "--warmup_epochs", "1", "--num_workers", "0",
        "--wandb_disabled", "--save_frequency", "1"]


def write_frames(root, n_train=8, n_val=2):
    """(64, 1024, 2) float32 range + intensity frames in metres, as kitti_utils/sample_kitti_dataset.py writes them."""
    rng = np.random.default_rng(0)
    for split, n in (("train", n_train), ("val", n_val)):
        os.makedirs(os.path.join(root, split), exist_ok=True)
        for i in range(n):
            a = np.zeros((64, 1024, 2), np.float32)
            a[..., 0] = rng.uniform(2.0, 80.0, (64, 1024)) * (rng.uniform(size=(64, 1024)) > 0.15)
            a[..., 1] = rng.uniform(0, 1, (64, 1024))
            np.save(os.path.join(root, split, f"{i:08d}.npy"), a)


def run_driver(extra, dropin, nproc=1, timeout=900, args=None):
    cmd = [sys.executable]
    if nproc > 1:
        cmd += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
                "--master-port", "29533"]
    cmd += ["main_lidar_upsampling.py"] + (ARGS if args is None else args) + extra
    return subprocess.run(cmd, cwd=cenv.REF_TULIP, env=cenv.driver_env(dropin), capture_output=True, text=True, timeout=timeout)


@needs_ref
def test_driver_imports_resolve_to_reference_or_dropin():
    probe = "import main_lidar_upsampling, engine_upsampling; import model.tulip as t; print(t.tulip_base.__module__)"
    for dropin, want in ((False, "model.tulip"), (True, "tulip_b200.model.tulip")):
        r = subprocess.run([sys.executable, "-c", probe], cwd=cenv.REF_TULIP, env=cenv.driver_env(dropin), capture_output=True,
                           text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        assert r.stdout.strip().splitlines()[-1] == want


@needs_ref
def test_layer_decay_groups_follow_named_parameter_chunks():
    import importlib.util
    spec = importlib.util.spec_from_file_location("_shim_optim_factory", os.path.join(cenv.SHIMS, "timm", "optim", "optim_factory.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    param_groups_layer_decay = mod.param_groups_layer_decay
    m = torch.nn.Sequential(*[torch.nn.Linear(4, 4) for _ in range(13)])           # 26 parameters -> 3 chunks of 12 + head slot
    groups = param_groups_layer_decay(m, 0.05)
    assert sum(len(g["params"]) for g in groups) == 26
    assert all((g["weight_decay"] == 0.0) == all(p.ndim == 1 for p in g["params"]) for g in groups)
    assert sorted({round(g["lr_scale"], 6) for g in groups}) == sorted({round(0.75 ** k, 6) for k in (1, 2, 3)})


@needs_ref
@pytest.mark.gpu
def test_unmodified_driver_trains_resumes_and_evaluates_on_dropin(tmp_path):
    data, out = str(tmp_path / "KITTI"), str(tmp_path / "out")
    write_frames(data)
    common = ["--data_path_low_res", data, "--data_path_high_res", data, "--output_dir", out, "--log_dir", out]
    r = run_driver(common + ["--epochs", "2"], dropin=True)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "Training finished" in r.stdout
    ckpts = sorted(glob.glob(os.path.join(out, "checkpoint-*.pth")))
    assert [os.path.basename(c) for c in ckpts] == ["checkpoint-0.pth", "checkpoint-1.pth"]
    log = [json.loads(l) for l in open(os.path.join(out, "log.txt"))]
    assert len(log) == 2 and all(np.isfinite(e["train_loss"]) for e in log)
    ck = torch.load(ckpts[-1], map_location="cpu", weights_only=False)
    assert len(ck["model"]) == 226 and set(ck) >= {"model", "optimizer", "epoch", "scaler", "args"}
    # resume through misc.load_model (util/misc.py:361-382, strict load_state_dict) and train one more epoch
    r2 = run_driver(common + ["--epochs", "3", "--resume", ckpts[-1]], dropin=True)
    assert r2.returncode == 0, (r2.stdout[-3000:], r2.stderr[-3000:])
    assert "Resume checkpoint" in r2.stdout and os.path.exists(os.path.join(out, "checkpoint-2.pth"))
    # evaluate() (engine_upsampling.py:126-355): forward under autocast, post-processing, point clouds, Chamfer shim, voxel metrics
    r3 = run_driver(common + ["--eval"], dropin=True)
    assert r3.returncode == 0, (r3.stdout[-3000:], r3.stderr[-3000:])
    assert os.path.exists(os.path.join(out, "results.txt"))


@needs_ref
@pytest.mark.gpu
def test_unmodified_driver_trains_the_expanding_variants(tmp_path):
    """The same unmodified driver WITHOUT --pixel_shuffle / --patch_unmerging: PatchExpanding in the decoder and the
    FinalPatchExpanding head (tulip.py:126-159) -- the configuration no shipped script selects."""
    data, out = str(tmp_path / "KITTI"), str(tmp_path / "out")
    write_frames(data, n_train=4, n_val=2)
    common = ["--data_path_low_res", data, "--data_path_high_res", data, "--output_dir", out, "--log_dir", out]
    args = [a for a in ARGS if a not in ("--pixel_shuffle", "--patch_unmerging")]
    r = run_driver(common + ["--epochs", "1"], dropin=True, args=args)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    ck = torch.load(os.path.join(out, "checkpoint-0.pth"), map_location="cpu", weights_only=False)
    keys = set(ck["model"])
    assert {"final_patch_expanding.expand.weight", "final_patch_expanding.norm.bias", "first_patch_expanding.norm.weight",
            "layers_up.0.upsample.expand.weight", "layers_up.1.upsample.norm.bias"} <= keys
    assert not any(k.startswith("ps_head") or k.endswith("upsample.expand.bias") for k in keys)
    log = [json.loads(l) for l in open(os.path.join(out, "log.txt"))]
    assert len(log) == 1 and np.isfinite(log[0]["train_loss"])


@needs_ref
@pytest.mark.gpu
def test_unmodified_driver_under_torchrun_two_ranks(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (DDP wraps the drop-in module, main_lidar_upsampling.py:276-278)")
    data, out = str(tmp_path / "KITTI"), str(tmp_path / "out")
    write_frames(data)
    common = ["--data_path_low_res", data, "--data_path_high_res", data, "--output_dir", out, "--log_dir", out]
    r = run_driver(common + ["--epochs", "1"], dropin=True, nproc=2)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert os.path.exists(os.path.join(out, "checkpoint-0.pth"))
