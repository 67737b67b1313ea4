"""A/B of the two-kernel classify path on ONE box: python tools/split_ab.py lib[:split] ...  (lib = in-tree or a path; split = 1 / 0).
Each spec runs in its own process: equality with the fused kernel's outputs on 65,536 clips, best-of-3 step time, and the two kernels' times."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import os, sys
sys.path.insert(0, %r)
import torch, eikws_pkg
m = eikws_pkg.load()
n = int(os.environ.get("AB_N", "65536"))
imp = m.Impulse(os.environ.get("AB_MODEL", "l476"))
split = os.environ.get("AB_SPLIT", "1") == "1"
clips = imp.synth_clips_device(n)
imp.set_split(False)
ref = imp.run_classifier_device(clips).clone()
imp.set_split(split)
out = torch.empty((n, imp.label_count), dtype=torch.float32, device="cuda:0")
for _ in range(3):
    imp.run_classifier_device(clips, out=out)
torch.cuda.synchronize()
same = torch.equal(ref, out)
best = 1e9
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(8):
        imp.run_classifier_device(clips, out=out)
    b.record()
    torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b) / 8)
k = ""
if split:
    imp.set_kernel_timing(True)
    for _ in range(5):
        imp.run_classifier_device(clips, out=out)
    ka, kb, _ = imp.split_kernel_ms()
    k = "spectral %%.3f ms  cepstral %%.3f ms" %% (ka, kb)
print("%%-22s split=%%d  %%.3f ms/step  %%7.3f M clips/s  equal=%%s  %%s" %% (os.path.basename(os.environ.get("EIKWS_B200_LIB", "in-tree")), split, best, n / best / 1e3, same, k), flush=True)
""" % ROOT

for rnd in range(int(os.environ.get("AB_ROUNDS", "2"))):
    for spec in sys.argv[1:]:
        lib, split = (spec.split(":") + ["1"])[:2]
        env = dict(os.environ)
        env["AB_SPLIT"] = split
        if lib != "in-tree":
            env["EIKWS_B200_LIB"] = os.path.abspath(lib)
        subprocess.run([sys.executable, "-c", CHILD], env=env, check=False)
