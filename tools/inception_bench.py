"""InceptionV3 inference throughput (BASELINE config 5: 299x299x3 uint8, B=128, end point Mixed_7c).
    python tools/inception_bench.py [--batch 128] [--precision mixed] [--reps 5]"""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imitation_from_observation_b200 import inception as I, ops      # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--precision", default="mixed")
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
net = I.InceptionV3(precision=a.precision, seed=0)
frames = torch.randint(0, 256, (a.batch, 299, 299, 3), dtype=torch.uint8, device="cuda")
for _ in range(2):
    net.forward(frames, ("Mixed_7c",))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    net.forward(frames, ("Mixed_7c",))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
macs = 0
H = 299
print("InceptionV3 %s B=%d: %.2f ms / batch, %.0f img/s, %.1f TFLOP/s (11.42 GFLOP/img)" %
      (a.precision, a.batch, ms, a.batch / ms * 1e3, a.batch * 11.42e9 / ms / 1e9))
ops.PROFILE = []
shapes = []
_orig = ops.conv2d_general
def _wrap(x, w, *a_, **k_):
    shapes.append((tuple(x.shape), tuple(w.shape), k_.get("stride", 1), k_.get("padding", "SAME")))
    return _orig(x, w, *a_, **k_)
ops.conv2d_general = _wrap
net.forward(frames, ("Mixed_7c",))
ops.conv2d_general = _orig
torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
agg = {}
for kind, impl, flops, s, e in prof:
    k = {1: "simt", 2: "tc1", 3: "tc3"}.get(impl, "?")
    d = agg.setdefault(k, [0, 0.0, 0.0]); d[0] += 1; d[1] += flops; d[2] += s.elapsed_time(e)
for k, (n, f, t) in agg.items():
    print("  %-5s %3d conv/gemm launches  %7.2f ms  %6.1f TFLOP/s  (%.0f%% of conv flops)" % (k, n, t, f / t / 1e9, 100 * f / (a.batch * 11.42e9)))

rows = []
conv_prof = [p for p in prof if p[0] == "fprop"]
for (xs, ws, st, pad), (kind, impl, flops, s_, e_) in zip(shapes, conv_prof):
    rows.append((s_.elapsed_time(e_), flops, xs, ws, st, pad))
rows.sort(reverse=True)
print("slowest convs:")
for t, f, xs, ws, st, pad in rows[:16]:
    print("  %7.3f ms %6.1f TF/s  x%s w%s s%d %s" % (t, f / t / 1e9, xs[1:], ws, st, pad))
