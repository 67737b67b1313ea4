"""Touch every kernel variant once with small batches -- the command compute-sanitizer is pointed at:
   compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_run.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import eikws_pkg

m = eikws_pkg.load()
synth = sys.modules["eikws_b200.synth"] if "eikws_b200.synth" in sys.modules else __import__("eikws_b200.synth", fromlist=["x"])
n = 613  # odd, a little more than one wave of clip groups
clips = synth.synth_clips(n, first_clip=5)
d16 = torch.from_numpy(clips).to("cuda:0")
d32 = (d16.to(torch.float32) / 32768.0).contiguous()
quick = "--quick" in sys.argv  # racecheck: the default model only, every schedule / lowering knob
# every lowering of the fused int8 path on the default model: mode 6 (default), 5, 4, 7 (two clip groups), 7 (one), 2
imp = m.Impulse("l476")
want = imp.run_classifier_device(d16).clone()  # the two-kernel path (default for int16 clips)
for m_clips in (1, 7, 8, 9, 163):  # work units that straddle clips / reach beyond the batch, CTAs with one clip
    assert torch.equal(imp.run_classifier_device(d16[:m_clips].contiguous()), want[:m_clips]), m_clips
imp.set_split(False)  # the fused kernel in every lowering
assert torch.equal(imp.run_classifier_device(d16), want)
for knobs in ({"work_claiming": False}, {"cmvn_shortcut": False}, {"tensor_core": False}, {"tensor_core": False, "clips_per_cta": 1},
              {"tensor_core": False, "cmvn_shortcut": False}):
    imp.set_work_claiming(knobs.get("work_claiming", True))
    imp.set_cmvn_shortcut(knobs.get("cmvn_shortcut", True))
    imp.set_tensor_core(knobs.get("tensor_core", True))
    imp.set_clips_per_cta(knobs.get("clips_per_cta", 2))
    got = imp.run_classifier_device(d16)
    torch.cuda.synchronize()
    assert torch.equal(got, want), knobs
imp.close()
for name in (("l476", "zip6") if quick else ("l476", "l432", "gsc12", "dw3", "zip6", "l476f32")):
    imp = m.Impulse(name)
    p16 = imp.run_classifier_device(d16)
    p32 = imp.run_classifier_device(d32)
    feats = imp.extract_mfcc_features_device(d16)
    pinf = imp.run_inference_device(feats)
    torch.cuda.synchronize()
    assert torch.equal(p16, p32) and torch.equal(p16, pinf), name
    mfe = imp.extract_mfe_features_device(d16)
    mfe32 = imp.extract_mfe_features_device(d32)
    torch.cuda.synchronize()
    assert torch.equal(torch.nan_to_num(mfe), torch.nan_to_num(mfe32))
    host = imp.run_classifier(clips[:40])
    assert np.array_equal(host, p16[:40].cpu().numpy())
    if name in ("l476", "l432", "gsc12", "dw3", "zip6"):
        st = m.Streams(imp, 7)
        audio = clips[:14].reshape(7, -1)
        for s in range(8):
            st.push(audio[:, s * st.slice_size:(s + 1) * st.slice_size])
        st.close()
    imp.close()
imp = m.Impulse("l476")
# tests-only CMVN stage kernel (both paths) and the multi-device host entry
cep = np.random.default_rng(3).standard_normal((97, 49, 13)).astype(np.float32) * 5
assert np.array_equal(imp.debug_cmvn_quantise(cep, True), imp.debug_cmvn_quantise(cep, False))
multi = m.MultiImpulse("l476")
assert np.array_equal(multi.run_classifier(clips[:101]), imp.run_classifier(clips[:101]))
multi.close()
i2s = torch.randint(-2 ** 31, 2 ** 31 - 1, (4 * 16000 * 3,), dtype=torch.int32, device="cuda:0")
imp.decimate_i2s_device(i2s, 16000 * 3)
torch.cuda.synchronize()
print("all kernel variants ran")
