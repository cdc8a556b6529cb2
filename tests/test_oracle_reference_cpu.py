"""Live differential check of the oracle against the unmodified reference on random model shapes, batches and masks
(build container only: skipped when the reference tree is not mounted; subprocess so that `model.aread` can resolve
to the reference there)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("AREAD_REF", "/root/reference")


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "model", "aread.py")), reason="reference tree not mounted")
def test_oracle_matches_the_reference_on_random_configurations():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_oracle_differential.py")], cwd=ROOT,
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "ORACLE DIFFERENTIAL OK" in res.stdout, res.stdout[-1500:] + res.stderr[-3000:]
