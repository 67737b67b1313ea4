"""Small driver for ncu captures: runs the fused kernel a few times over a device-resident synthetic batch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import eikws_pkg

m = eikws_pkg.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
imp = m.Impulse(os.environ.get("EIKWS_MODEL", "l476"))
if os.environ.get("EIKWS_TC", "") != "":  # 0 = dp4a block 1, 1 = UMMA, 2 = UMMA + run-ahead schedule
    imp.set_tensor_core(int(os.environ["EIKWS_TC"]))
if os.environ.get("EIKWS_PIPE", "") != "":  # 1 = the software-pipelined kernel
    imp.set_pipelined(os.environ["EIKWS_PIPE"] == "1")
if os.environ.get("EIKWS_SPLIT", "") != "":  # 0 = the fused kernel, 1 = the two-kernel path (default)
    imp.set_split(os.environ["EIKWS_SPLIT"] == "1")
clips = imp.synth_clips_device(n)
if os.environ.get("EIKWS_F32", "") == "1":  # float32 samples (x / 32768, exact)
    clips = (clips.to(torch.float32) / 32768.0).contiguous()
out = torch.empty((n, imp.label_count), dtype=torch.float32, device="cuda:0")
for _ in range(reps):
    imp.run_classifier_device(clips, out=out)
torch.cuda.synchronize()
print("done", out[:2].tolist())
