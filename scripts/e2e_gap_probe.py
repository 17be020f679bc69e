"""Host time between the loss read-back of step i and the forward launch of step i+1 (the part of the end-to-end step during which
the GPU idles), split by what the host does in between."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tulip_b200.model import tulip as T
from tulip_b200.model.tulip import tulip_base
B = 32
torch.manual_seed(0)
model = tulip_base(img_size=(16, 1024), target_img_size=(64, 1024), patch_size=(1, 4), in_chans=1, window_size=[2, 8], swin_v2=False,
                   pixel_shuffle=True, circular_padding=True, log_transform=True, patch_unmerging=True).cuda().train()
lo = torch.rand(B, 1, 16, 1024, device="cuda"); hi = torch.rand(B, 1, 64, 1024, device="cuda")
marks = {}
orig_launch = model._launch_forward
def launch(*a, **k):
    marks["pre_launch"] = time.perf_counter()
    r = orig_launch(*a, **k)
    marks["post_launch"] = time.perf_counter()
    return r
model._launch_forward = launch
orig_sample = model._sample_drop_scales
def sample(*a, **k):
    marks["pre_sample"] = time.perf_counter()
    r = orig_sample(*a, **k)
    marks["post_sample"] = time.perf_counter()
    return r
model._sample_drop_scales = sample
import tulip_b200._lib as L
lib = L.load_library()
class Timed:
    def __init__(self, fn, name): self.fn, self.name = fn, name
    def __call__(self, *a):
        t = time.perf_counter(); r = self.fn(*a); marks[self.name] = marks.get(self.name, 0.0) + time.perf_counter() - t; return r
for nm in ("tulip_net_forward", "tulip_net_backward", "tulip_net_backward_phases"):
    if hasattr(lib, nm):
        setattr(lib, nm, Timed(getattr(lib, nm), nm))
acc = {}
def add(k, v): acc[k] = acc.get(k, 0.0) + v
N = 40
for it in range(N + 5):
    t_start = time.perf_counter()
    model.zero_grad(set_to_none=True)
    t_zg = time.perf_counter()
    _, loss, _ = model(lo, hi)
    t_fwd = time.perf_counter()
    loss.backward()
    t_bwd = time.perf_counter()
    v = loss.item()
    t_item = time.perf_counter()
    if it >= 5:
        add("zero_grad", t_zg - t_start)
        add("forward: entry -> drop-scale sampling", marks["pre_sample"] - t_zg)
        add("forward: drop-scale sampling (4 ATen ops)", marks["post_sample"] - marks["pre_sample"])
        add("forward: sampling -> _launch_forward", marks["pre_launch"] - marks["post_sample"])
        add("forward: _launch_forward (3 copies + graph launch)", marks["post_launch"] - marks["pre_launch"])
        add("forward: rest (autograd Function.apply)", t_fwd - marks["post_launch"])
        add("backward call", t_bwd - t_fwd)
        add("item() wait", t_item - t_bwd)
        add("whole step", t_item - t_start)
        add("  C call tulip_net_forward (graph launch)", marks.get("tulip_net_forward", 0.0))
        add("  C call tulip_net_backward (graph launch)", marks.get("tulip_net_backward", 0.0) + marks.get("tulip_net_backward_phases", 0.0))
    marks["tulip_net_forward"] = marks["tulip_net_backward"] = marks["tulip_net_backward_phases"] = 0.0
for k, v in acc.items():
    print(f"{k:55s} {v / N * 1e3:8.3f} ms")
print("pre-drawn DropPath sets used:", getattr(model, "_predraw_hits", 0), "of", N + 5)
