"""Stage-by-stage comparison of the CUDA kernel against the plain-C oracle (debug aid; run on the GPU box)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import eikws_pkg

m = eikws_pkg.load()
import eikws_b200.synth as synth
from oracle_lib import PortOracle

name = sys.argv[1] if len(sys.argv) > 1 else "l476"
imp = m.Impulse(name)
port = PortOracle(name)
clips = np.concatenate([synth.synth_clips(8), np.stack(list(synth.special_clips().values()))])
names = [f"synth{i}" for i in range(8)] + list(synth.special_clips().keys())
lib = m.load_library()
n = clips.shape[0]
rec = C.c_int(0)
lib.eikws_debug_stage_taps_i16_host(None, None, 0, None, C.byref(rec))
taps = np.zeros((n, rec.value), np.float32)
rc = lib.eikws_debug_stage_taps_i16_host(imp._h, C.c_void_p(clips.ctypes.data), n, C.c_void_p(taps.ctypes.data), None)
assert rc == 0, lib.eikws_last_error()
feats_o, otaps = port.mfcc_i16(clips, taps=True)
probs, feats, q = imp.run_classifier_taps(clips)


def cmp(label, a, b):
    bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
    msg = f"  {label:8s} mismatches {int(bad.sum()):6d}/{a.size}"
    if bad.any():
        idx = np.argwhere(bad)[:4]
        msg += "  first: " + "; ".join(f"{tuple(int(v) for v in i)} gpu={a[tuple(i)]!r} ref={b[tuple(i)]!r}" for i in idx)
    print(msg)


for i in range(n):
    P = taps[i, :129 * 49].reshape(129, 49).T                      # [49][129]
    L = taps[i, 129 * 49:129 * 49 + 49 * 33].reshape(49, 33)[:, :32]
    F = taps[i, 129 * 49 + 49 * 33:].reshape(49, 13)
    print(names[i])
    cmp("power", P, otaps[i]["power"])
    en = otaps[i]["energy"]
    with np.errstate(all="ignore"):
        cmp("mfcc", F, otaps[i]["mfcc"])
    cmp("features", feats[i].reshape(49, 13), feats_o[i].reshape(49, 13))
    # log-mel: oracle tap is mel before the log; compare through the pre-CMVN cepstra instead, and show raw L stats
    print("   logmel finite:", bool(np.isfinite(L).all()))
