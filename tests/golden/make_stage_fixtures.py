#!/usr/bin/env python
"""Builds tests/golden/{instancing,nested_instancing,curves}.stage.json: the reference's sample stages as parsed by
this repo's text-USD reader (crust-render_b200/usd/usda.py), serialised with usda.stage_to_json. scenes.sample_stage()
re-imports them through usd/importer.py, so the native-instancing / nested-instancer / BasisCurves samples run on the
GPU box, where /root/reference does not exist. tests/test_usd_import.py checks that each fixture still equals a fresh
parse of the file (where the file exists)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from crust_render_b200.usd import usda  # noqa: E402

SAMPLES = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/samples"
HERE = os.path.dirname(os.path.abspath(__file__))
for name in ("instancing", "nested_instancing", "curves"):
    doc = usda.stage_to_json(usda.load(os.path.join(SAMPLES, name + ".usda")))
    dst = os.path.join(HERE, name + ".stage.json")
    with open(dst, "w") as f:
        json.dump(doc, f, separators=(",", ":"))
    print(dst, os.path.getsize(dst), "bytes")
