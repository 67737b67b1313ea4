"""A/B timing of kernel builds on ONE box: python tools/ab.py libA.so libB.so[:clips_per_cta] ...  (each timed in its own process)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import os, sys
sys.path.insert(0, %r)
import torch, eikws_pkg
m = eikws_pkg.load()
n = 65536
imp = m.Impulse(os.environ.get("AB_MODEL", "l476"))
G = int(os.environ.get("AB_G", "1"))
if G != 1:
    small = imp.synth_clips_device(1001)
    ref = imp.run_classifier_device(small).clone()
    imp.set_clips_per_cta(G)
    assert torch.equal(ref, imp.run_classifier_device(small)), "clips_per_cta changes results"
if os.environ.get("AB_TC", "") != "":
    imp.set_tensor_core(os.environ["AB_TC"] == "1")
if os.environ.get("AB_CMVN", "") != "":
    imp.set_cmvn_shortcut(os.environ["AB_CMVN"] == "1")
if os.environ.get("AB_DYN", "") != "":
    imp.set_work_claiming(os.environ["AB_DYN"] == "1")
if os.environ.get("AB_PIPE", "") != "":
    imp.set_pipelined(os.environ["AB_PIPE"] == "1")
SKEW = os.environ.get("AB_SKEW", "")
if SKEW:
    imp.set_skew_ns(int(SKEW))
clips = imp.synth_clips_device(n)
out = torch.empty((n, imp.label_count), dtype=torch.float32, device="cuda:0")
for _ in range(3):
    imp.run_classifier_device(clips, out=out)
torch.cuda.synchronize()
best = 0.0
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(8):
        imp.run_classifier_device(clips, out=out)
    b.record()
    torch.cuda.synchronize()
    best = max(best, n * 8 / (a.elapsed_time(b) * 1e-3))
print("%%-24s skew=%%-6s G=%%d tc=%%-2s cmvn=%%-2s dyn=%%-2s pipe=%%-2s %%8.3f M clips/s  checksum %%.6f" %% (os.path.basename(os.environ.get("EIKWS_B200_LIB", "in-tree")), SKEW or "dflt", G, os.environ.get("AB_TC", "") or "-", os.environ.get("AB_CMVN", "") or "-", os.environ.get("AB_DYN", "") or "-", os.environ.get("AB_PIPE", "") or "-", best / 1e6, float(out.double().sum())), flush=True)
""" % ROOT

for rnd in range(int(os.environ.get('AB_ROUNDS', '2'))):
    for spec in sys.argv[1:]:
        lib, g, skew, tc, cm, dyn, pipe = (spec.split(":") + ["", "", "", "", "", ""])[:7]
        env = dict(os.environ)
        env["AB_G"] = g or "1"
        env["AB_SKEW"] = skew
        env["AB_TC"] = tc
        env["AB_CMVN"] = cm
        env["AB_DYN"] = dyn
        env["AB_PIPE"] = pipe
        if lib != "in-tree":
            env["EIKWS_B200_LIB"] = os.path.abspath(lib)
        subprocess.run([sys.executable, "-c", CHILD], env=env, check=False)
