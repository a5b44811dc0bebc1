"""CPU tier: the shipped C-ABI library loads and exports every symbol include/algames_b200.h declares; descriptor
sizes (no compute calls without a GPU); the package never imports the oracle."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as g
    g.build()
    return g.LIB


def test_library_exports_every_declared_symbol(built):
    header = open(os.path.join(ROOT, "include", "algames_b200.h")).read()
    declared = set(re.findall(r"\b(agb_[a-z0-9_]+)\s*\(", header))
    import algames_b200 as ab
    assert declared == set(ab._capi.SYMBOLS), declared ^ set(ab._capi.SYMBOLS)
    lib = C.CDLL(built)
    for name in declared:
        assert hasattr(lib, name), name


def test_abi_layout_check(built):
    """agb_abi_check: the hand-written ctypes mirrors agree with the header the library was built from, and a drifted
    binding (one more field in a struct, another AGB_NHIST) is refused with a message naming the entry."""
    import algames_b200 as ab
    lib = ab._capi.load()                       # load() itself runs the check
    words = ab._capi.abi_layout()
    mine = (C.c_int * len(words))()
    assert lib.agb_abi_layout(mine, len(words)) == 0 and list(mine) == words
    for k, name in ((0, "sizeof(agb_problem_desc)"), (13, "offsetof(agb_options, dual_reset)"), (24, "AGB_NHIST")):
        bad = list(words); bad[k] += 4
        assert lib.agb_abi_check((C.c_int * len(bad))(*bad), len(bad)) != 0
        assert name in lib.agb_last_error(None).decode()
    assert lib.agb_abi_check((C.c_int * 3)(1, 2, 3), 3) != 0


def test_sizes_of_descriptor(built):
    import algames_b200 as ab
    lib = ab._capi.load()
    for name, (n, m, p, N, S, nrow) in {"B": (12, 6, 3, 40, 2106, 6), "C": (16, 8, 4, 50, 4312, 28), "A": (8, 4, 2, 20, 532, 16)}.items():
        cfg = ab.workloads.CONFIGS[name]() if name == "A" else ab.workloads.CONFIGS[name](batch=1)
        d = ab.problem._make_desc(cfg[0], cfg[1], cfg[2], cfg[3], cfg[4])
        sz = ab._capi.Sizes()
        assert lib.agb_sizes_of(C.byref(d), C.byref(sz)) == 0
        assert (sz.n, sz.m, sz.p, sz.N, sz.S, sz.nrow) == (n, m, p, N, S, nrow)


def test_default_options_match_reference(built):
    import algames_b200 as ab
    lib = ab._capi.load()
    o = ab._capi.OptionsC()
    lib.agb_default_options(C.byref(o))
    ref = ab.Options().to_c()
    for f, _ in ab._capi.OptionsC._fields_:
        a, b = getattr(o, f), getattr(ref, f)
        assert (list(a) == list(b)) if hasattr(a, "__len__") else (a == b), f


def test_no_device_fails_loudly(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import algames_b200 as ab
    cfg = ab.workloads.config_a()
    with pytest.raises(ab.AlgamesError, match="no CUDA device"):
        ab.GameBatch(cfg[0], cfg[1], cfg[2], cfg[3], cfg[4], 1)


def test_package_does_not_import_oracle():
    code = "import sys; sys.path.insert(0, %r); import algames_b200; assert not [m for m in sys.modules if m.startswith('oracle')]" % ROOT
    subprocess.run([sys.executable, "-c", code], check=True)
