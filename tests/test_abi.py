"""CPU tests of the drop-in boundary: libb200pc.so loads, exports every symbol include/b200pc.h
declares (and nothing the header does not), argument validation works without a GPU, and the
product never imports the oracle."""
import ctypes as C
import os
import re
import subprocess

import pytest

from b200pc import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200pc.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200pc_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_in_tree():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build()"
    assert _lib.LIB_PATH.startswith(ROOT)


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    decl = declared_symbols()
    assert len(decl) >= 18
    for name in decl:
        assert hasattr(lib, name), "header declares %s but the library does not export it" % name
        assert name in _lib.SIGNATURES, "%s has no ctypes prototype in b200pc/_lib.py" % name
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = sorted(set(re.findall(r"\sT\s+(b200pc_[a-z0-9_]+)", out)))
    assert exported == decl, "exported C symbols and header differ: %s" % (set(exported) ^ set(decl))


def test_library_targets_sm_100a_only():
    out = subprocess.check_output(["cuobjdump", "--list-elf", _lib.LIB_PATH], text=True)
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_hot_loop_uses_packed_fp32_and_bulk_tma():
    sass = subprocess.check_output(["cuobjdump", "-sass", _lib.LIB_PATH], text=True)
    for mnemonic in ("FFMA2", "FMUL2", "FADD2", "FMNMX3", "UBLKCP", "SYNCS"):
        assert mnemonic in sass, "expected %s in the SASS of libb200pc.so" % mnemonic


def _sass_by_function():
    sass = subprocess.check_output(["cuobjdump", "-sass", _lib.LIB_PATH], text=True)
    out, name = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            out[name] = []
        elif name is not None:
            out[name].append(line)
    return out


def test_fps_kernels_never_fuse_multiply_add():
    """farthest_point_sample computes (dx*dx + dy*dy) + dz*dz with every product rounded on its own
    (Utils/Pointnet2Utils.py:80).  ptxas 12.9 contracts packed mul+add into FFMA2, so the kernel spells the
    products with scalar .rn intrinsics; any fused multiply-add in its SASS would flip near-tied picks."""
    funcs = {n: b for n, b in _sass_by_function().items() if "fps_kernel" in n or "fps_flat_kernel" in n}
    assert len(funcs) == 10, sorted(funcs)          # two kernel families x P = 1, 2, 4, 8, 16
    for name, body in funcs.items():
        text = "\n".join(body)
        for bad in ("FFMA2", "FMUL2", " FFMA ", "FFMA.", "DFMA"):
            assert bad not in text, "%s contains %s" % (name, bad.strip())
        assert " FMUL " in text and " FADD " in text


def test_torch_ops_mirror_the_header():
    """every compute entry of include/b200pc.h is a registered torch.ops.b200pc.<name> (CUDA key only, with a fake
    kernel), and no op exists without a C entry behind it"""
    import torch
    from b200pc import ops
    not_ops = {"last_error", "version", "device_sm_count", "fma_peak", "tuning_reload"}
    compute = {n[len("b200pc_"):] for n in declared_symbols()}
    compute = {n for n in compute if n not in not_ops and not n.endswith("_workspace_bytes") and not n.endswith("_host")}
    assert compute == set(ops.OP_SCHEMAS), compute ^ set(ops.OP_SCHEMAS)
    for name in compute:
        op = getattr(torch.ops.b200pc, name).default
        assert torch._C._dispatch_has_kernel_for_dispatch_key(op.name(), "CUDA"), name
        assert not torch._C._dispatch_has_kernel_for_dispatch_key(op.name(), "CPU"), "%s must not have a CPU kernel" % name
    with pytest.raises(NotImplementedError):        # the dispatcher itself refuses CPU tensors: no fallback
        torch.ops.b200pc.knn(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3), 2, 0, False)
    for name in ("gather", "group_points", "three_interpolate", "feature_propagation", "chamfer_fwd"):
        op = getattr(torch.ops.b200pc, name).default
        assert torch._C._dispatch_has_kernel_for_dispatch_key(op.name(), "Autograd"), name


def test_fake_kernels_infer_shapes_without_a_gpu():
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode
    from b200pc import ops  # noqa: F401
    with FakeTensorMode():
        ref = torch.empty(2, 100, 3, device="cuda"); qry = torch.empty(2, 50, 3, device="cuda")
        idx, dist = torch.ops.b200pc.knn(ref, qry, 4, 0, True)
        assert idx.shape == (2, 50, 4) and idx.dtype == torch.int64 and dist.shape == (2, 50, 4)
        assert torch.ops.b200pc.ball_query(ref, qry, 1.0, 8).shape == (2, 50, 8)
        feat = torch.empty(2, 100, 16, device="cuda")
        assert torch.ops.b200pc.group_points(ref, qry, feat, idx, True).shape == (2, 19, 4, 50)
        resi, nn, gf, i2 = torch.ops.b200pc.fusion_group(qry, ref, feat, 8)
        assert resi.shape == (2, 4, 50, 8) and nn.shape == (2, 3, 50, 8) and gf.shape == (2, 16, 50, 8) and i2.shape == (2, 50, 8)
        out, i3, w3 = torch.ops.b200pc.feature_propagation(ref, qry, torch.empty(2, 50, 32, device="cuda"), 0)
        assert out.shape == (2, 100, 32) and i3.shape == (2, 100, 3) and w3.shape == (2, 100, 3)


def test_version_and_workspace_queries_need_no_gpu():
    lib = _lib.load()
    assert lib.b200pc_version() >= 100
    small = lib.b200pc_search_workspace_bytes(1, 1024, 256, 16)
    big = lib.b200pc_search_workspace_bytes(8, 16384, 16384, 16)
    assert 8 * 16384 * 16 <= big < 64 * 2 ** 20
    assert 0 < small
    assert lib.b200pc_fps_workspace_bytes(1, 16384) > 0


def test_argument_validation_returns_einval_with_message():
    lib = _lib.load()
    null = C.c_void_p(0)
    rc = lib.b200pc_knn(null, null, 1, 10, 10, 3, 0, null, null, null, 0, null)
    assert rc == _lib.EINVAL and "null" in _lib.last_error()
    one = C.c_void_p(16)   # never dereferenced: validation fails first
    rc = lib.b200pc_knn(one, one, 1, 4, 2, 5, 0, one, null, one, 1 << 20, null)
    assert rc == _lib.EINVAL and "exceeds" in _lib.last_error()
    rc = lib.b200pc_knn(one, one, 1, 4096, 128, 4, 0, one, null, one, 16, null)
    assert rc == _lib.EWORKSPACE
    rc = lib.b200pc_three_nn(one, one, 1, 10, 2, 0, one, one, null, one, 1 << 20, null)
    assert rc == _lib.EINVAL and "3" in _lib.last_error()
    rc = lib.b200pc_fps(one, 1, 10 ** 7, 4, one, one, null, 0, null)
    assert rc == _lib.EINVAL


def test_no_gpu_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("box has a GPU")
    lib = _lib.load()
    rc = lib.b200pc_device_sm_count()
    assert rc == _lib.ECUDA
    from b200pc import pointnet2_utils as P
    with pytest.raises(RuntimeError):
        P.square_distance(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "point-cloud-interpolation-_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("the oracle side can use it too", ""), os.path.join(dirpath, f)
