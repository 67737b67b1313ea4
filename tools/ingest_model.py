#!/usr/bin/env python
"""Turn an UNMODIFIED Edge Impulse export (model-parameters/ + tflite-model/) into an EIKWSMDL container.

    python tools/ingest_model.py <export_root> <out.eikwsmdl>

<export_root> is the directory that holds `model-parameters/model_metadata.h` and
`tflite-model/trained_model_compiled.cpp` (e.g. the reference's
embedded-demos/stm32cubeide/nucleo-l476-keyword-spotting/ei-keyword-spotting).  Nothing is copied: the
generated files are compiled where they lie, against this repo's include/ shim of the TFLite C operator API,
and linked with libeikws_b200.so whose Register_*() operators record the graph (csrc/tflm_capture.cpp).
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "ei-keyword-spotting_b200")


def ingest(export_root: str, out_path: str) -> None:
    meta = os.path.join(export_root, "model-parameters", "model_metadata.h")
    model = os.path.join(export_root, "tflite-model", "trained_model_compiled.cpp")
    for p in (meta, model):
        if not os.path.isfile(p):
            raise FileNotFoundError(p)
    m = re.search(r"ei_dsp_config_mfcc_t\s+(ei_dsp_config_\w+)\s*=", open(meta).read())
    if not m:
        raise RuntimeError("no ei_dsp_config_mfcc_t instance in " + meta + " (not an MFCC impulse)")
    cfg = m.group(1)
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "ingest")
        cmd = [
            "g++", "-std=gnu++14", "-O1", "-w",
            "-I" + os.path.join(ROOT, "include"),  # must precede the export so OUR tflite shim headers win
            "-I" + export_root,
            "-DEIKWS_MFCC_CFG=" + cfg,
            os.path.join(ROOT, "tools", "ingest", "ingest_main.cpp"), model,
            "-L" + LIBDIR, "-leikws_b200", "-Wl,-rpath," + LIBDIR,
            "-o", exe,
        ]
        subprocess.run(cmd, check=True)
        subprocess.run([exe, out_path], check=True)


if __name__ == "__main__":
    if len(sys.argv) != 3:
        sys.exit(__doc__)
    ingest(sys.argv[1], sys.argv[2])
