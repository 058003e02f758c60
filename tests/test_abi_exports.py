"""The product library (nvcc build of libpgpd.so, not the emulator) loads without a GPU and exports every entry point
that include/pgpd.h declares; host-only entry points answer.  No kernel is launched here."""
import ctypes as C
import os
import re

from pointnetgpd_b200 import _abi as A
from pointnetgpd_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "pgpd.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pgpd_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_the_declared_abi():
    path = B.build()
    assert os.path.basename(path) == "libpgpd.so"
    lib = C.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), "libpgpd.so does not export %s declared in include/pgpd.h" % name
    # the Python binding declares the same set
    assert set(A.EXPORTS) <= set(declared), set(A.EXPORTS) - set(declared)


def test_host_only_entry_points_answer_without_a_gpu():
    lib = A.bind(C.CDLL(B.build()))
    assert lib.pgpd_version() == 100
    small = lib.pgpd_workspace_bytes(A.PGPD_CLS, 2, 64, 2, 0)
    train = lib.pgpd_workspace_bytes(A.PGPD_CLS, 2, 64, 2, A.F_TRAIN | A.F_SAVE)
    big = lib.pgpd_workspace_bytes(A.PGPD_CLS, 512, 1024, 2, A.F_TRAIN | A.F_SAVE)
    assert 0 < small < train < big < 8 << 30          # B=512, N=1024 training fits in a few GB of the 180 GB
    assert lib.pgpd_workspace_bytes(A.PGPD_CLS, 0, 64, 2, 0) == 0
    assert lib.pgpd_tower_workspace_bytes(4, 100, A.F_TRAIN | A.F_SAVE) > 0
