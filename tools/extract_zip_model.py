"""Unpack the generated model files (model-parameters/, tflite-model/) of an Edge Impulse Arduino library zip.

usage: extract_zip_model.py <library.zip> <out_dir>

The reference repo ships its third model only inside embedded-demos/arduino/.../ei-keyword-spotting-03-arduino-1.0.2.zip
(6 labels; conv k3 -> pool 2 -> conv k3 -> pool 2 -> FC 208 -> 6).  The files are written
UNMODIFIED into a git-ignored build directory (oracle/_ref/gen/...), from where the reference harness and
tools/ingest_model.py compile them exactly like the two STM32 exports; nothing from the zip enters the repository.
"""
import os
import sys
import zipfile


def extract(zip_path: str, out_dir: str) -> None:
    with zipfile.ZipFile(zip_path) as z:
        for name in z.namelist():
            if name.endswith("/"):
                continue
            parts = name.split("/")
            if len(parts) >= 3 and parts[0] == "src" and parts[1] in ("model-parameters", "tflite-model"):
                dst = os.path.join(out_dir, *parts[1:])
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                with open(dst, "wb") as f:
                    f.write(z.read(name))


if __name__ == "__main__":
    extract(sys.argv[1], sys.argv[2])
