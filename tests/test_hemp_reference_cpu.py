"""Live differential check of the HEMP host logic against the unmodified reference (build container only: skipped
when the reference tree is not mounted).  Runs in a subprocess so that `model.aread` can resolve to the reference
there while this process keeps the repository's drop-in modules."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("AREAD_REF", "/root/reference")


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "model", "aread.py")), reason="reference tree not mounted")
def test_validate_and_generate_match_the_reference_on_random_masks():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_hemp_differential.py")], cwd=ROOT,
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "HEMP DIFFERENTIAL OK" in res.stdout, res.stdout[-1500:] + res.stderr[-3000:]
