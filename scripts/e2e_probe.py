"""Where does the end-to-end gap come from?  Times variants of the step loop (CUDA events around 30 steps) and the host
time of the forward / backward calls."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tulip_b200.model.tulip import tulip_base
B = 32
torch.manual_seed(0)
model = tulip_base(img_size=(16, 1024), target_img_size=(64, 1024), patch_size=(1, 4), in_chans=1, window_size=[2, 8], swin_v2=False,
                   pixel_shuffle=True, circular_padding=True, log_transform=True, patch_unmerging=True).cuda().train()
lo = torch.rand(B, 1, 16, 1024, device="cuda"); hi = torch.rand(B, 1, 64, 1024, device="cuda")

def step():
    model.zero_grad(set_to_none=True)
    _, loss, _ = model(lo, hi)
    loss.backward()
    return loss

def timed(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

print("device loop (no sync)        %.3f ms" % timed(step))
print("sync every step (.item())    %.3f ms" % timed(lambda: step().item()))
prev = [None]
def lagged():
    l = step()
    if prev[0] is not None: prev[0].item()
    prev[0] = l
print("lagged .item() (prev step)   %.3f ms" % timed(lagged))
# host time of the two calls
torch.cuda.synchronize()
tf = tb = 0.0
for _ in range(10):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); model.zero_grad(set_to_none=True); _, loss, _ = model(lo, hi); t1 = time.perf_counter()
    loss.backward(); t2 = time.perf_counter()
    tf += t1 - t0; tb += t2 - t1
print("host time per step: forward call %.3f ms, backward call %.3f ms" % (tf * 100, tb * 100))
import cProfile, pstats, io
pr = cProfile.Profile()
torch.cuda.synchronize()
pr.enable()
for _ in range(20):
    step().item()
pr.disable()
sio = io.StringIO(); pstats.Stats(pr, stream=sio).sort_stats("cumulative").print_stats(22); print(sio.getvalue()[:4500])
