"""Where does an iteration of eikws_cepstral_kernel go?  Needs a library built with -DEIKWS_CEP_TRACE=1 (tools/build_variant.sh trace
-DEIKWS_CEP_TRACE=1; EIKWS_B200_LIB=ab/libeikws_trace.so python tools/cep_trace.py): per warp role, mean cycles per iteration of the
stage-1 task, the wait at the first CTA barrier, the CMVN, the wait at the last CTA barrier."""
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
for _ in range(2):
    probs, q = imp.run_classifier_taps_device(clips)
torch.cuda.synchronize()
ctas = 148 * 6
tr = q.view(-1)[: ctas * 5 * 8 * 8].view(torch.int64).view(ctas, 5, 8).double().cpu()
it = tr[:, :, 4].clamp(min=1)
names = ["stage-1 task", "wait at barrier 1", "CMVN (+ resolutions)", "wait at last barrier"]
roles = ["warp 0 epilogue + block 2", "warp 1 epilogue + block 2", "warp 2 DCT rows 0-31", "warp 3 DCT rows 32-48", "warp 4 tail"]
print("mean cycles per iteration (over %d CTAs, %.0f iterations each)" % (ctas, float(it.mean())))
for w in range(5):
    vals = [(tr[:, w, i] / it[:, w]).mean().item() for i in range(4)]
    print("  %-28s" % roles[w], "  ".join("%s %7.0f" % (names[i], vals[i]) for i in range(4)), "  sum %7.0f" % sum(vals))
loop = tr[:, :, 5]
per_it = (loop / it)
q = torch.quantile(per_it.flatten(), torch.tensor([0.0, 0.1, 0.5, 0.9, 1.0], dtype=torch.float64))
print("cycles of the whole clip loop per iteration, over CTAs x warps: min %.0f  p10 %.0f  median %.0f  p90 %.0f  max %.0f" % tuple(q.tolist()))
tot = torch.quantile(loop[:, 0], torch.tensor([0.0, 0.5, 1.0], dtype=torch.float64))
print("cycles of the whole clip loop per CTA: min %.0f  median %.0f  max %.0f   (kernel: 0.705 ms = %.0f cycles at 1965 MHz)" % (*tot.tolist(), 0.705e-3 * 1.965e9))
print("(BAR.SYNC is deferred-blocking: a clock read right after a barrier is taken BEFORE the warp blocks, so each column holds the wait of the barrier before it)")
