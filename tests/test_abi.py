"""The C-ABI library loads and exports every symbol include/*.h declares; the ctypes structs match
the header field for field.  No compute calls (runs without a GPU)."""
import ctypes
import glob
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = importlib.import_module("aread-multi-domain-recommendation_b200._lib")
build = importlib.import_module("aread-multi-domain-recommendation_b200.build")


def declared_functions():
    names = []
    for header in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(header).read(), flags=re.S)
        names += re.findall(r"AREAD_API\s+[\w\s\*]+?\b(aread_\w+)\s*\(", text)
    return sorted(set(names))


def struct_fields(name):
    text = "".join(open(h).read() for h in glob.glob(os.path.join(ROOT, "include", "*.h")))
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        # "float momentum, eps, dropout_p" declares several fields
        decl = re.sub(r"\[[^\]]*\]", "", decl)                  # array extents: `int32_t dims[4][3]` declares `dims`
        names = [re.search(r"(\w+)\s*$", part.strip()).group(1) for part in decl.split(",")]
        fields.extend(names)
    return fields


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    names = declared_functions()
    assert len(names) >= 6
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert set(_lib.exported_symbols()) == set(names), "ctypes binding and header disagree"
    assert lib.aread_abi_version() == _lib.ABI_VERSION
    assert lib.aread_last_error() == b""


def declared_structs():
    text = "".join(open(h).read() for h in glob.glob(os.path.join(ROOT, "include", "*.h")))
    return sorted(set(re.findall(r"typedef struct (aread_\w+) \{", text)))


def binding_name(cname):
    """aread_grouped_linear_args -> GroupedLinearArgs (the naming rule of _lib.py)."""
    return "".join(part.capitalize() for part in cname[len("aread_"):].split("_"))


@pytest.mark.parametrize("cname", declared_structs())
def test_struct_fields_match_header(cname):
    ctype = getattr(_lib, binding_name(cname), None)
    assert ctype is not None, f"{cname} is declared in include/ but _lib.py has no {binding_name(cname)}"
    fields = [f.rstrip("_") for f, _ in ctype._fields_]      # `in` is a Python keyword
    assert fields == struct_fields(cname)


def test_every_header_struct_has_a_binding():
    declared = declared_structs()
    assert len(declared) >= 19
    for cname in declared:
        assert hasattr(_lib, binding_name(cname)), cname


def test_struct_sizes_are_native(lib):
    assert ctypes.sizeof(_lib.EmbedPlan) == 56
    assert ctypes.sizeof(_lib.GatherArgs) == 56 + 8 * 9
    assert ctypes.sizeof(_lib.ScatterArgs) == 56 + 8 * 12


def test_argument_validation_without_gpu(lib):
    """Bad arguments are rejected before any CUDA call is made."""
    args = _lib.GatherArgs()
    assert lib.aread_gather_fwd(ctypes.byref(args), None) == _lib.AREAD_ERR_INVALID
    assert b"plan" in lib.aread_last_error()
    with pytest.raises(_lib.AreadError):
        _lib.check(lib.aread_scatter_bwd(None, None))


def test_hei_path_rules_need_no_gpu():
    """Which tower-layer implementation a shape gets is host logic (include/aread_sm100.h, aread_hei_layer_path):
    bit 0 = forward on the tensor cores, bit 1 = backward."""
    lib = _lib.load()
    lib.aread_hei_set_path(-1, -1)
    default = lib.aread_hei_layer_path(65536, 3, 64, 64)
    assert default in (0, 3), "the environment switches both directions unless AREAD_HEI_TC_BWD=0"
    lib.aread_hei_set_path(1, 1)
    try:
        assert lib.aread_hei_layer_path(65536, 3, 64, 64) == 3 and lib.aread_hei_layer_path(65536, 12, 16, 8) == 3
        assert lib.aread_hei_layer_path(65536, 6, 32, 16) == 3 and lib.aread_hei_layer_path(512, 1, 16, 16) == 3
        assert lib.aread_hei_layer_path(511, 3, 64, 64) == 0, "under 512 rows: CUDA cores"
        assert lib.aread_hei_layer_path(65536, 4, 20, 10) == 0 and lib.aread_hei_layer_path(65536, 4, 64, 8) == 0
        lib.aread_hei_set_path(1, 0)
        assert lib.aread_hei_layer_path(65536, 3, 64, 64) == 1
        lib.aread_hei_set_path(0, 0)
        assert lib.aread_hei_layer_path(65536, 3, 64, 64) == 0
    finally:
        lib.aread_hei_set_path(-1, -1)
    assert lib.aread_rowpass_prologue_ctas(65536) >= 1 and lib.aread_rowpass_prologue_ctas(1) == 1
