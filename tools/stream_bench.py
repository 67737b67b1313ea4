"""Throughput of continuous mode (run_classifier_continuous over many lock-step streams): python tools/stream_bench.py [n_streams] [pushes]
Every push delivers one 250 ms slice (4000 int16 samples) per stream; once the window is full each push classifies every stream."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import eikws_pkg

m = eikws_pkg.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
pushes = int(sys.argv[2]) if len(sys.argv) > 2 else 16
imp = m.Impulse("l476")
clips = imp.synth_clips_device(2 * n)  # 8 slices of audio per stream
ref = None
for shortcut in (False, True):
    imp.set_cmvn_shortcut(shortcut)
    st = m.Streams(imp, n)
    probs = torch.zeros((n, imp.label_count), dtype=torch.float32, device="cuda:0")
    flat = clips.view(-1)
    sl = [flat[k * n * st.slice_size:(k + 1) * n * st.slice_size].view(n, st.slice_size).contiguous() for k in range(8)]
    for k in range(8):  # fill the window, warm up
        st.push_device(sl[k % 8], probs)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(pushes):
        st.push_device(sl[k % 8], probs)
    b.record()
    torch.cuda.synchronize()
    t = a.elapsed_time(b) * 1e-3
    if ref is None:
        ref = probs.clone()
    else:
        assert torch.equal(ref, probs), "the shortcut changes continuous-mode outputs"
    print(f"continuous mode, {n} streams, cmvn shortcut {'on ' if shortcut else 'off'}: {n * pushes / t / 1e6:7.3f} M stream-slices/s "
          f"= {n * pushes * 0.25 / t / 1e6:6.3f} M audio-seconds/s  ({t / pushes * 1e3:.3f} ms per push)", flush=True)
    st.close()
