#!/usr/bin/env python
"""Recipe: ship the UNMODIFIED reference modules of the hot path to the GPU box. TEST / BASELINE INFRASTRUCTURE ONLY.

/root/reference exists in the dev container only. This copies — verbatim, no edits — the few Python files the image path
imports into ``oracle/_ref/reference/`` with their directory layout, plus the licence. ``oracle/_ref/``
is git-ignored (the repository history never holds reference sources) but NOT gpurun-ignored, so it travels with the
snapshot like the built .so files; there ``oracle/ref_harness.py`` imports it for
  * ``bench.py --impl reference`` (the reference's own PyTorch code on the box's host cores),
  * the GPU-eager reference leg (the same code on the B200 under CUDA autocast: SURVEY.md section 8d's "number to beat"),
  * ``-m gpu`` tests that pin the oracle's autocast-bf16 mode against the reference running under the real CUDA autocast.
``__graft_entry__.build()`` runs this whenever /root/reference is present.

  python oracle/make_ref.py
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

SRC = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "reference")
FILES = [
    "LICENSE",
    "modeling/t2i_pipeline.py",
    "modeling/utils.py",
    "modeling/vision_head/flow_head_parallel_x.py",
    "modeling/vision_head/sampling_x.py",
    "modeling/vision_encoder/autoencoder.py",
    "utils/fs.py",
    "imagenet_gen/src/model_parallel.py",
    "imagenet_gen/src/layers_parallel.py",
    "imagenet_gen/src/diff_head_parallel.py",
    "imagenet_gen/src/sampling_parallel.py",
    "imagenet_gen/src/gfq.py",
    "imagenet_gen/src/qae.py",
]


def make(verbose: bool = False) -> str | None:
    if not os.path.isdir(os.path.join(SRC, "modeling")):
        return DST if os.path.isdir(DST) else None
    manifest = {}
    for rel in FILES:
        s = os.path.join(SRC, rel)
        if not os.path.exists(s):
            continue
        d = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest[rel] = hashlib.sha256(open(d, "rb").read()).hexdigest()
    # every other module imagenet_gen/src imports (copied whole: small, pure Python)
    src_dir = os.path.join(SRC, "imagenet_gen", "src")
    if os.path.isdir(src_dir):
        for f in sorted(os.listdir(src_dir)):
            if f.endswith(".py") and f"imagenet_gen/src/{f}" not in manifest:
                d = os.path.join(DST, "imagenet_gen", "src", f)
                os.makedirs(os.path.dirname(d), exist_ok=True)
                shutil.copyfile(os.path.join(src_dir, f), d)
                manifest[f"imagenet_gen/src/{f}"] = hashlib.sha256(open(d, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1)
    if verbose:
        print(f"{len(manifest)} reference files -> {DST}")
    return DST


if __name__ == "__main__":
    make(verbose=True)
