"""The namespace-package drop-in: with this repository ahead of the reference on sys.path, `model.aread`
and `model.layer` resolve here and every other `model.*` module to the reference.  Needs the reference
tree, so it only runs in the build container."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("AREAD_REF", "/root/reference")


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "model", "aread.py")), reason="reference tree not mounted")
def test_namespace_shadowing():
    code = f"""
import sys
sys.path[:0] = [{ROOT!r}, {REF!r}]
import model.aread, model.layer, model.dfm, config
assert model.aread.__file__.startswith({ROOT!r}), model.aread.__file__
assert model.layer.__file__.startswith({ROOT!r}), model.layer.__file__
assert model.dfm.__file__.startswith({REF!r}), model.dfm.__file__
assert model.aread.AREAD.__module__.endswith("_b200.aread")
from model.layer import FactorizationMachine, DNN, CrossNetV2      # baseline-only layers: served from the reference
assert FactorizationMachine.__module__ == "_aread_reference_layer"
print("ok")
"""
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]
