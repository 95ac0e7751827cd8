"""The C-ABI library loads on a GPU-less box and exports every symbol include/mppib.h declares; the ctypes
struct mirrors have the same size as the C structs (checked against a tiny gcc-compiled probe)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT
from mppi_isaac_b200 import backend
from mppi_isaac_b200.model import blob


def _declared():
    src = open(os.path.join(ROOT, "include", "mppib.h")).read()
    return sorted(set(re.findall(r"\b(mppib_[a-z_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(backend.lib_path()):
        import __graft_entry__
        __graft_entry__.build()
    lib = C.CDLL(backend.lib_path())
    names = _declared()
    assert set(names) == set(backend.SYMBOLS)
    for n in names:
        assert hasattr(lib, n), n
    lib.mppib_abi_version.restype = C.c_int32
    assert lib.mppib_abi_version() == blob.ABI_VERSION


def test_struct_layout_matches_header(tmp_path):
    probe = tmp_path / "probe.c"
    probe.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "mppib.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(MppibModel), sizeof(MppibParams),'
                     ' offsetof(MppibModel, link_body), offsetof(MppibModel, contact_kp), offsetof(MppibParams, obs));return 0;}\n')
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(probe), "-o", str(exe)])
    out = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert out == [C.sizeof(blob.MppibModel), C.sizeof(blob.MppibParams), blob.MppibModel.link_body.offset,
                   blob.MppibModel.contact_kp.offset, blob.MppibParams.obs.offset]


def test_no_cpu_fallback():
    """The product path refuses to run without CUDA instead of silently computing on the host."""
    import torch
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        backend.CudaBackend("cpu")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            backend.CudaBackend("cuda:0")
