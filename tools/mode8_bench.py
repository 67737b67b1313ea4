"""generic int8 op plan (zip6) with and without the certified CMVN shortcut: python tools/mode8_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, eikws_pkg
m = eikws_pkg.load()
imp = m.Impulse("zip6")
n = 32768
d = imp.synth_clips_device(n)
out = torch.empty((n, imp.label_count), dtype=torch.float32, device="cuda:0")
ref = None
for sc in (False, True):
    imp.set_cmvn_shortcut(sc)
    for _ in range(3): imp.run_classifier_device(d, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(8): imp.run_classifier_device(d, out=out)
    b.record(); torch.cuda.synchronize()
    if ref is None: ref = out.clone()
    else: assert torch.equal(ref, out), "the shortcut changes the generic plan's outputs"
    print(f"zip6 (generic int8 op plan) shortcut {'on ' if sc else 'off'}: {n*8/(a.elapsed_time(b)*1e-3)/1e6:7.3f} M clips/s", flush=True)
