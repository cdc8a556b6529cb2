"""Structure rules of the repository: the product never touches the oracle, the reference tree or a CPU path, and the
drop-in modules stay a namespace package."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "aread-multi-domain-recommendation_b200")


def _sources(folder, exts):
    for base, _, files in os.walk(folder):
        if "_build" in base or "__pycache__" in base:
            continue
        for f in files:
            if f.endswith(exts):
                yield os.path.join(base, f)


def test_product_does_not_import_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    offenders = [p for p in list(_sources(PKG, (".py",))) + list(_sources(os.path.join(ROOT, "model"), (".py",)))
                 if pat.search(open(p).read())]
    assert not offenders, offenders


def test_only_layer_getattr_mentions_the_reference_tree():
    # baseline-only symbols of model.layer are served lazily from a reachable reference tree (layer.py::__getattr__);
    # nothing else in the product may depend on /root/reference
    offenders = [p for p in _sources(PKG, (".py", ".cu", ".cuh")) if "/root/reference" in open(p).read()
                 and not p.endswith("layer.py")]
    assert not offenders, offenders


def test_model_is_a_namespace_package():
    assert not os.path.exists(os.path.join(ROOT, "model", "__init__.py"))
    assert sorted(f for f in os.listdir(os.path.join(ROOT, "model")) if f.endswith(".py")) == ["aread.py", "layer.py"]


def test_no_forbidden_batch_copy_calls():
    names = ("MemcpyBatch" + "Async", "Memcpy3DBatch" + "Async")
    offenders = [p for p in _sources(ROOT, (".py", ".cu", ".cuh", ".h", ".md", ".sh"))
                 if any(n in open(p, errors="ignore").read() for n in names)]
    assert not offenders, offenders
