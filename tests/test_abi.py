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


def _mapping(model, monkeypatch=None, **env):
    import copy
    lib = backend.load_library()
    old = {k: os.environ.get(k) for k in ("MPPIB_K2_LANES", "MPPIB_K2_TEAM")}
    try:
        for k in old:
            os.environ.pop(k, None)
        os.environ.update(env)
        return backend.CudaBackend.MAPPING_NAMES[lib.mppib_rollout_mapping_for_model(C.byref(model))].split("-")[0]
    finally:
        for k, v in old.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v


def test_which_kernel_takes_which_scene():
    """mppib_rollout_mapping_for_model (host arithmetic, no GPU): serial chains without contacts -> lanes kernel; trees and every
    contact scene within the team kernel's limits (<= 16 bodies numbered depth first) -> team kernel; anything else, and the A/B
    knobs, -> thread-per-rollout"""
    import copy
    from scenes import boxer_setup, gripper_setup, panda_setup, push_setup, robot_setup
    panda, gripper = panda_setup(K=8, T=4)[0].model, gripper_setup(K=8, T=4)[0].model
    push, boxer = push_setup(K=8, T=4)[0].model, boxer_setup(K=8, T=4)[0].model
    albert = robot_setup("albert", "mmrobot_link7", K=8, T=4)[0].model
    assert _mapping(panda) == "lanes"
    assert [_mapping(m) for m in (gripper, push, boxer, albert)] == ["team"] * 4
    # the A/B knobs
    assert _mapping(panda, MPPIB_K2_LANES="0") == "team"                          # a chain is a tree
    assert _mapping(panda, MPPIB_K2_LANES="0", MPPIB_K2_TEAM="0") == "thread"
    assert _mapping(gripper, MPPIB_K2_TEAM="0") == "thread" and _mapping(push, MPPIB_K2_TEAM="0") == "thread"
    assert _mapping(panda, MPPIB_K2_TEAM="1") == "lanes"                          # the lanes kernel keeps the chains it can take
    # outside the team kernel's limits: bodies not numbered depth first (the two fingers hang below body 6; renumber one of them
    # BEFORE a body of the arm's chain), and a chain with contacts stays eligible while a contact-free one goes to the lanes kernel
    bad = copy.deepcopy(gripper)
    par = [bad.parent[i] for i in range(bad.nb)]
    assert par == [-1, 0, 1, 2, 3, 4, 5, 6, 6]
    for i, pv in enumerate([-1, 0, 1, 2, 3, 3, 4, 6, 6]):                          # body 5 now hangs below body 3, next to body 4's subtree [4, 6, ...]
        bad.parent[i] = pv
    assert _mapping(bad) == "thread"
    lib = backend.load_library()
    assert lib.mppib_rollout_mapping_for_model(None) < 0
