import os, sys
sys.path.insert(0, '/root/repo')
import torch, eikws_pkg
m = eikws_pkg.load()
imp = m.Impulse("l476")
n = 32768
d = imp.synth_clips_device(n)
x = (d.to(torch.float32) / 32768.0).contiguous()
out = torch.empty((n, imp.label_count), dtype=torch.float32, device="cuda:0")
for label, clips, tc in (("float-input", x, True), ("int16 dp4a (tensor core off)", d, False)):
    imp.set_tensor_core(tc)
    for sc in (False, True):
        imp.set_cmvn_shortcut(sc)
        for _ in range(3): imp.run_classifier_device(clips, out=out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(8): imp.run_classifier_device(clips, out=out)
        b.record(); torch.cuda.synchronize()
        print(f"{label:30s} shortcut {'on ' if sc else 'off'}: {n*8/(a.elapsed_time(b)*1e-3)/1e6:7.3f} M clips/s", flush=True)
