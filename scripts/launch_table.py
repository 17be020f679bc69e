"""Per-launch table of one training step (executor's built-in profiler: CUDA events around every launch).
usage: python scripts/launch_table.py [batch] > gpurun_out/launch_table.md"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tulip_b200._lib import load_library
from tulip_b200.model.tulip import tulip_base

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
lib = load_library()
torch.manual_seed(0)
model = tulip_base(img_size=(16, 1024), target_img_size=(64, 1024), patch_size=(1, 4), in_chans=1, window_size=[2, 8], swin_v2=False,
                   pixel_shuffle=True, circular_padding=True, log_transform=True, patch_unmerging=True).cuda().train()
lo = torch.rand(B, 1, 16, 1024, device="cuda")
hi = torch.rand(B, 1, 64, 1024, device="cuda")


def step():
    model.zero_grad(set_to_none=True)
    pred, loss, pix = model(lo, hi)
    loss.backward()


for _ in range(4):
    step()
torch.cuda.synchronize()
NREP = 3
per = None
for rep in range(NREP):
    lib.tulip_net_profile(model._net, 1)
    step()
    torch.cuda.synchronize()
    n = lib.tulip_net_profile_record(model._net, -1, None, None, None, None)
    rows = []
    tag, ms, fl, by = C.c_int(), C.c_double(), C.c_double(), C.c_double()
    for i in range(n):
        lib.tulip_net_profile_record(model._net, i, C.byref(tag), C.byref(ms), C.byref(fl), C.byref(by))
        rows.append([tag.value, ms.value, fl.value, by.value])
    if per is None:
        per = rows
    else:
        for a, b in zip(per, rows):
            a[1] = min(a[1], b[1])
lib.tulip_net_profile(model._net, 0)
name = C.create_string_buffer(64)
d = C.c_double(); q = C.c_int64()
names = {}
for t in range(lib.tulip_net_profile_num_tags()):
    lib.tulip_net_profile_read(model._net, t, name, 64, C.byref(d), C.byref(d), C.byref(d), C.byref(q))
    names[t] = name.value.decode()
print(f"# per-launch table, batch {B}, min of {NREP} steps; total {sum(r[1] for r in per):.3f} ms over {len(per)} launches\n")
print("| # | kernel | us | GFLOP | MB | TFLOP/s | GB/s |\n|---:|---|---:|---:|---:|---:|---:|")
for i, (t, ms_, fl_, by_) in enumerate(per):
    print(f"| {i} | {names[t]} | {ms_ * 1e3:.1f} | {fl_ / 1e9:.2f} | {by_ / 1e6:.1f} | {fl_ / ms_ / 1e9 if ms_ else 0:.0f} | {by_ / ms_ / 1e6 if ms_ else 0:.0f} |")
