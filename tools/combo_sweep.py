"""Throughput of every (model, input sample type) combination on one GPU: python tools/combo_sweep.py [n_clips]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import eikws_pkg

m = eikws_pkg.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
for name in ("l476", "l432", "gsc12", "dw3", "zip6", "l476f32"):
    imp = m.Impulse(name)
    pcm = imp.synth_clips_device(n)
    for dtype in ("int16", "float32"):
        clips = pcm if dtype == "int16" else (pcm.to(torch.float32) / 32768.0).contiguous()
        out = torch.empty((n, imp.label_count), dtype=torch.float32, device="cuda:0")
        for _ in range(2):
            imp.run_classifier_device(clips, out=out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            imp.run_classifier_device(clips, out=out)
        b.record()
        torch.cuda.synchronize()
        print(f"{name:8s} {dtype:8s} {n * 5 / (a.elapsed_time(b) * 1e-3) / 1e6:7.3f} M clips/s", flush=True)
        del clips
    del pcm

# MFCC DSP block alone (extract_mfcc_features): 32,000 B in, 2,548 B out per clip
imp = m.Impulse("l476")
pcm = imp.synth_clips_device(n)
feat = imp.extract_mfcc_features_device(pcm)
for _ in range(2):
    imp.extract_mfcc_features_device(pcm, features=feat)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    imp.extract_mfcc_features_device(pcm, features=feat)
b.record()
torch.cuda.synchronize()
print(f"MFCC block int16 {n * 5 / (a.elapsed_time(b) * 1e-3) / 1e6:7.3f} M clips/s", flush=True)
del pcm, feat

# the sibling MFE DSP block (features only: 32,000 B in, 6,272 B out per clip)
imp = m.Impulse("l432")
pcm = imp.synth_clips_device(n)
feat = imp.extract_mfe_features_device(pcm)
for _ in range(2):
    imp.extract_mfe_features_device(pcm, features=feat)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    imp.extract_mfe_features_device(pcm, features=feat)
b.record()
torch.cuda.synchronize()
thr = n * 5 / (a.elapsed_time(b) * 1e-3)
print(f"MFE block int16 {thr / 1e6:7.3f} M clips/s = {thr * (32000 + 6272) / 1e9:6.1f} GB/s algorithmic", flush=True)
