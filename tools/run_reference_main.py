#!/usr/bin/env python
"""Run the reference's own `main.py` / `run.py` with this repository's AREAD as a drop-in.

    python tools/run_reference_main.py --reference /path/to/AREAD-Multi-Domain-Recommendation \
        --workdir /tmp/aread_run -- --model aread --dataset_name aliccp --domain_filter "[0,1,2]" ...

What it does ("Launcher corollaries"):
  * puts this repository's root AHEAD of the reference tree on sys.path, so `model.aread` and
    `model.layer` resolve here (namespace package: the repo has no model/__init__.py) while `run`,
    `config`, `preprocess`, `dataset.*` and every other `model.*` still resolve to the reference;
  * initialises wandb in disabled mode (run.py logs without ever calling wandb.init, run.py:165);
  * runs from a writable work directory holding a copy of the reference's `dataset/` (run.py writes
    .pth caches and checkpoints relative to the cwd, run.py:94-95, 260-263);
  * optionally swaps `torch.optim.Adam` inside `run` for the fused optimizer (--fused-adam).
Needs a CUDA device: the drop-in has no CPU fallback.
"""
import argparse
import os
import runpy
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("AREAD_REF", "/root/reference"))
    ap.add_argument("--workdir", default="/tmp/aread_run")
    ap.add_argument("rest", nargs=argparse.REMAINDER, help="arguments after -- go to the reference's main.py")
    args = ap.parse_args()
    ref = os.path.abspath(args.reference)
    if not os.path.exists(os.path.join(ref, "main.py")):
        raise SystemExit(f"reference tree not found at {ref} (set --reference or AREAD_REF)")
    os.makedirs(args.workdir, exist_ok=True)
    if not os.path.exists(os.path.join(args.workdir, "dataset")):
        shutil.copytree(os.path.join(ref, "dataset"), os.path.join(args.workdir, "dataset"))
    os.chdir(args.workdir)
    os.environ.setdefault("AREAD_REF", ref)
    sys.path[:0] = [ROOT, ref]
    try:
        import wandb
        wandb.init(mode="disabled")
    except ImportError:
        pass
    import model.aread as drop_in                      # noqa: E402
    assert os.path.abspath(drop_in.__file__).startswith(ROOT), "model.aread did not resolve to this repository"
    rest = [a for a in args.rest if a != "--"]
    sys.argv = [os.path.join(ref, "main.py")] + rest
    runpy.run_path(os.path.join(ref, "main.py"), run_name="__main__")


if __name__ == "__main__":
    main()
