#!/usr/bin/env python
"""Golden record of the reference trainer's call sequence (run.py:578-686) executed by the UNMODIFIED reference
AREAD on CPU (build container only).

    python tests/golden/make_trainer_golden.py      ->  tests/golden/trainer_seq.pt

The sequence itself is tests/_trainer_sequence.run_sequence; the GPU test replays it on the CUDA module and
compares losses, candidate masks, pruned masks, eval losses and the selected masks.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.golden import make_golden as G            # noqa: E402  (puts the reference on sys.path when asked)
from tests import _trainer_sequence as T             # noqa: E402


def main():
    refcfg, AREAD = G.load_reference()
    assert "/reference/" in sys.modules["model.aread"].__file__ or os.environ.get("AREAD_REF")
    spec = G.Spec(**T.SPEC)
    model = G.build_reference(refcfg, AREAD, spec, dropout=0.0)
    out = T.run_sequence(model, torch.device("cpu"), spec)
    out["spec"], out["cfg"] = T.SPEC, T.SEQ
    path = os.path.join(HERE, "trainer_seq.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")
    print("warm-up losses", [round(v, 5) for v in out["warm_up"]])
    print("post losses", [round(v, 5) for v in out["post"]])
    print("selected masks active edges", [sum(sum(map(sum, m)) for m in mk) for mk in out["selected"]])


if __name__ == "__main__":
    main()
