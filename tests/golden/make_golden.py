#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/liboracle_ref_<model>.so, built by
`make -C oracle ref` from /root/reference).  Run in the build container only; the GPU box has no reference and
uses the committed files.  The reference ships no golden vectors of its own, so these are the
pinned known answers for every layer of the test pyramid.

Inputs are not stored: they are regenerated from (seed, index) by ei-keyword-spotting_b200/synth.py; of the hand-built special
clips the names are recorded, of the speech-like clips (synth.speechlike_clip: integer-only harmonic stacks with formants, a
syllable envelope and a noise burst, picked per model so that every label it can produce wins) the parameter rows.  Stored per model:
  features   [n,637] float32  extract_mfcc_features output (bit pattern matters)
  probs      [n,L]   float32  run_classifier output
  mel0/energy0/mfcc0          stage taps of clip 0 (mfe output, frame energies, pre-CMVN cepstra)
  filterbank [129,32]
  nn_features [m,637], nn_probs [m,L], nn_t19/21/23/25/27/29/30: run_inference on crafted feature vectors and the
  TFLite tensors that are still intact in the arena after invoke (see tests/test_oracle.py for the byte ranges)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import eikws_pkg  # noqa: E402

eikws_pkg.load()
import eikws_b200.synth as synth  # noqa: E402
from oracle_lib import RefOracle  # noqa: E402

N_SYNTH = 40
GOLDEN_SEED = 0xE1D5
N_SPEECH_CANDIDATES = 1500  # speech-like clips searched per model so that every label it can produce wins in the goldens
# arena tensors (index: byte range) that no later node overwrites, i.e. that can still be read after invoke;
# plus the two output-sized tensors at the end of every graph
INTACT = {19: (210, 1470), 21: (0, 210), 23: (0, 70), 25: (10, 70), 27: (0, 10)}     # L476 topology (l476, l432, gsc12, l476f32)
INTACT_BY_MODEL = {"zip6": {17: (160, 392), 25: (208, 400), 27: (0, 208)},           # Arduino-zip topology (arena offsets of its generated file)
                   # depthwise variant: its arena is planned without aliasing, every tensor survives
                   "dw3": {17: (0, 1470), 19: (0, 1470), 21: (0, 210), 23: (0, 210), 25: (0, 210), 27: (0, 30)}}


def crafted_features(n_labels_seed: int) -> np.ndarray:
    rng = np.random.default_rng(1234 + n_labels_seed)
    F = rng.normal(0, 1.5, (96, 637)).astype(np.float32)
    F[24:48] *= 4                                                      # saturates int8
    F[48:64] = rng.uniform(-300, 300, (16, 637)).astype(np.float32)    # wraps the float->int8 cast
    odd = np.array([1e6, -1e6, 1.4e7, -1.4e7, 2e9, -2e9, 5e9, -5e9, 1e20, -1e20, np.inf, -np.inf, np.nan, 0.0, -0.0], np.float32)
    F[64:72] = rng.choice(odd, (8, 637))
    F[72:96] = rng.normal(0, 0.3, (24, 637)).astype(np.float32)
    return F


def speech_candidates():
    """parameter rows of synth.speechlike_clip: (f0, glide, vowel a, vowel b, onset ms, length ms, level, burst ms, burst level, seed)"""
    rng = np.random.default_rng(42)
    n = N_SPEECH_CANDIDATES
    return np.stack([rng.integers(85, 260, n), rng.integers(-60, 80, n), rng.integers(0, 8, n), rng.integers(0, 8, n), rng.integers(50, 500, n),
                     rng.integers(150, 700, n), rng.integers(2000, 30000, n), rng.integers(0, 120, n), rng.integers(0, 6000, n),
                     rng.integers(0, 1 << 30, n)], 1).astype(np.int64)


def pick_speech_clips(ref, params, cand_clips):
    """for every label that wins on some candidate: the three candidates where it wins with the highest and the one where it wins
    with the lowest probability (a near-tie); for a model on which a single label always wins, six spread-out candidates"""
    probs = ref.run_classifier_i16(cand_clips)
    win = probs.argmax(1)
    chosen = []
    for lab in range(ref.n_labels):
        idx = np.nonzero(win == lab)[0]
        if idx.size == 0:
            continue
        order = idx[np.argsort(-probs[idx, lab])]
        chosen += list(order[:3]) + [order[-1]]
    if len(set(win)) <= 1:
        chosen += list(np.argsort(probs.max(1))[:: max(1, len(probs) // 6)][:6])
    chosen = sorted(set(int(c) for c in chosen))
    return params[chosen]


def main():
    cand_params = speech_candidates()
    cand_clips = synth.speechlike_clips(cand_params)
    for mi, name in enumerate(("l476", "l432", "gsc12", "l476f32", "zip6", "dw3")):
        ref = RefOracle(name)
        specials = synth.special_clips()
        speech_params = pick_speech_clips(ref, cand_params, cand_clips)
        clips = np.concatenate([synth.synth_clips(N_SYNTH, 0, GOLDEN_SEED), np.stack(list(specials.values())), synth.speechlike_clips(speech_params)])
        feats = ref.mfcc_i16(clips)
        probs = ref.run_classifier_i16(clips)
        mel0, en0 = ref.mfe_i16(clips[0])
        mfcc0 = ref.mfcc_nocmvn_i16(clips[0])
        # float-input path (config 5 style signal): same clips scaled as the demo callback would
        xf = (clips[:8].astype(np.float32) / np.float32(32768))
        feats_f32 = ref.mfcc_f32(xf)
        F = crafted_features(mi)
        nn_probs, tens = ref.run_inference(F, want_tensors=True)
        out = dict(n_synth=N_SYNTH, seed=GOLDEN_SEED, special_names=np.array(list(specials.keys())), speech_params=speech_params,
                   features=feats, probs=probs, mel0=mel0, energy0=en0, mfcc0=mfcc0, filterbank=ref.filterbank(),
                   features_f32in=feats_f32, nn_features=F, nn_probs=nn_probs, labels=np.array(ref.labels))
        n_t = len(tens[0])
        if ref.has_mfe_block:  # the sibling MFE DSP block exists only in the newer SDK copy (L432 build)
            out["mfe_features"] = ref.mfe_block_i16(clips)
        intact = INTACT_BY_MODEL.get(name, INTACT)
        out["intact"] = np.array([[k, lo, hi] for k, (lo, hi) in intact.items()], np.int32)
        for k, (lo, hi) in intact.items():
            out[f"nn_t{k}"] = np.stack([t[k][lo:hi] for t in tens])
        out["nn_t_fc"] = np.stack([t[n_t - 2] for t in tens])
        out["nn_t_out"] = np.stack([t[n_t - 1] for t in tens])
        path = os.path.join(HERE, f"golden_{name}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path), "bytes;", len(clips), "clips,", len(np.unique(probs, axis=0)), "distinct probability rows; argmax histogram",
              np.bincount(probs.argmax(1), minlength=ref.n_labels))


if __name__ == "__main__":
    main()
