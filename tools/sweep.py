"""Tuning sweep on the GPU box: CTAs per SM x start skew, device-resident 65,536-clip batch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import eikws_pkg

m = eikws_pkg.load()
n = 65536
imp = m.Impulse("l476")
clips = imp.synth_clips_device(n)
out = torch.empty((n, imp.label_count), dtype=torch.float32, device="cuda:0")


def run(steps=6):
    for _ in range(2):
        imp.run_classifier_device(clips, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        imp.run_classifier_device(clips, out=out)
    b.record()
    torch.cuda.synchronize()
    return n * steps / (a.elapsed_time(b) * 1e-3)


ref = None
for ctas in (3, 4):
    imp.set_ctas_per_sm(ctas)
    for skew in (0, 4000, 8000, 12000, 16000, 20000, 28000):
        imp.set_skew_ns(skew)
        v = run()
        if ref is None:
            ref = out.clone()
        assert torch.equal(ref, out)
        print(f"ctas/SM {ctas} skew {skew:6d} ns : {v / 1e6:7.3f} M clips/s", flush=True)
