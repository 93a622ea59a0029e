"""The C-ABI shared library loads (no GPU needed) and exports every symbol include/osufusion_b200.h declares."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_library_builds_loads_and_exports_header_symbols():
    from osufusion_b200 import _native
    from osufusion_b200.build import build_native

    lib_path = build_native()
    lib = ctypes.CDLL(str(lib_path))
    header = (ROOT / "include" / "osufusion_b200.h").read_text()
    declared = set(re.findall(r"\b(of_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    lib.of_version.restype = ctypes.c_int
    assert lib.of_version() >= 1


def test_binding_argument_counts_match_header():
    from osufusion_b200 import _native

    header = (ROOT / "include" / "osufusion_b200.h").read_text()
    for name, sig in _native._SIGS.items():
        m = re.search(r"int " + name + r"\((.*?)\);", header, re.S)
        assert m, name
        assert len(m.group(1).split(",")) == len(sig), name


def test_no_cpu_fallback_in_product_path():
    """The product package never imports the oracle, and refuses to run on CPU tensors."""
    import torch

    for f in (ROOT / "osufusion_b200").rglob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f
    from oracle.synth import TINY, synth_inputs
    from osufusion_b200.modules import UNet

    net = UNet(6, 96, 5, **TINY)
    x, a, c, t, _, _ = synth_inputs(1, 32, 1)
    try:
        net(x, a, t, c)
    except RuntimeError as e:
        assert "CUDA" in str(e)
    else:
        raise AssertionError("CPU call must fail loudly")


def test_state_dict_layout_matches_oracle():
    from oracle.denoiser import UNet as OracleUNet
    from oracle.synth import TINY
    from osufusion_b200.modules import UNet

    a, b = UNet(6, 96, 5, **TINY).state_dict(), OracleUNet(6, 96, 5, **TINY).state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)


def test_host_side_helpers_run_without_a_gpu():
    """Entry points that only touch host state: SM limit for the data-parallel overlap, grouped-kernel sizing helpers."""
    from osufusion_b200 import _native
    lib = _native.lib()
    assert lib.of_set_sm_limit(132) == 0 and lib.of_set_sm_limit(0) == 0
    assert lib.of_film_chunk_rows() > 0
    assert lib.of_pack_seg_ctas(512, 512, 3, 512) == 512 and lib.of_pack_seg_ctas(1024, 512, 1, 512) == 128
    assert lib.of_pack_seg_ctas(8, 8, 15, 8) == -1            # wide kernels use the per-tensor path
    assert lib.of_opt_tensor_ctas(4097) == 2


def test_argument_validation_returns_error_codes_without_a_gpu():
    """Every entry point validates its arguments before touching CUDA: a bad call returns OF_ERR_INVALID (-1) and of_last_error()
    names the entry point — the error behaviour of the C-ABI (include/osufusion_b200.h), checkable on the GPU-less box."""
    import ctypes as C

    from osufusion_b200 import _native as N
    lib = N.lib()

    def last():
        return lib.of_last_error().decode()

    g = N.GemmArgs()
    assert lib.of_gemm(None, None) == -1 and "of_gemm" in last()
    g.mode, g.batch, g.rows, g.N, g.K, g.taps = 0, 1, 16, 12, 8, 1            # N not a multiple of 8
    assert lib.of_gemm(C.byref(g), None) == -1 and "multiple of 8" in last()
    a = N.AttnArgs()
    a.B, a.H, a.KVH, a.L, a.D = 1, 4, 3, 16, 64                              # null pointers
    assert lib.of_attn_fwd(C.byref(a), None) == -1 and "of_attn_fwd" in last()
    buf = C.create_string_buffer(4096)
    p = C.addressof(buf)
    a.q = a.k = a.v = a.out = p
    a.D = 72                                                                 # head dim out of range
    assert lib.of_attn_fwd(C.byref(a), None) == -1 and "head dim" in last()
    a.D = 64                                                                 # 4 query heads on 3 kv heads
    assert lib.of_attn_fwd(C.byref(a), None) == -1 and "head counts" in last()
    assert lib.of_layernorm_fwd(p, 100, 4, 100, p, p, 1e-5, p, None, 100, None, None) == -1 and "of_layernorm_fwd" in last()
    assert lib.of_headnorm_fwd(p, 64, 64, 1, 1, 1, 1, 1, 72, p, p, 8.0, p, 64, 64, 0, None) == -1 and "head dim" in last()
    assert lib.of_headnorm_fwd(p, 64, 64, 1, 1, 1, 1, 1, 24, p, p, 4.9, p, 64, 64, 2, None) == -1 and "variant" in last()
    assert lib.of_adaln_fwd(p, 12, 12, 1, 1, 12, p, 12, p, 12, 1e-6, p, 12, 12, p, None) == -1 and "of_adaln_fwd" in last()
    assert lib.of_gate_bwd(p, 8, 8, p, 8, None, 8, 8, 1, 1, 1, 8, p, 8, 8, p, 8, None) == -1 and "null" in last()
    assert lib.of_linear_small_fwd(p, 8, 17, 8, 8, p, 8, None, 0, 1, p, 8, None, None) == -1 and "out of range" in last()
    assert lib.of_row_mean_std(p, 1, 1, 1, p, None) == -1 and "of_row_mean_std" in last()


def test_struct_layouts_match_the_ctypes_mirrors(tmp_path):
    """The C-ABI passes arguments in structs: include/osufusion_b200.h is compiled with gcc (plain C: the header must stay C-clean) and
    sizeof / offsetof of every field are compared with the ctypes mirrors in osufusion_b200/_native.py — a field added on one side only
    would otherwise corrupt arguments silently."""
    import ctypes as C
    import re
    import shutil
    import subprocess
    from pathlib import Path

    from osufusion_b200 import _native as N
    if shutil.which("gcc") is None:
        import pytest
        pytest.skip("gcc not available")
    root = Path(__file__).resolve().parent.parent
    header = (root / "include" / "osufusion_b200.h").read_text()
    pairs = {"of_gemm_args": N.GemmArgs, "of_attn_args": N.AttnArgs, "of_rb_args": N.RbArgs, "of_film_group": N.FilmGroup,
             "of_pack_seg": N.PackSeg, "of_lora_finish_seg": N.LoraFinishSeg, "of_opt_tensor": N.OptTensor}
    declared = set(re.findall(r"^\} (of_\w+);", header, flags=re.M))
    assert declared == set(pairs), declared ^ set(pairs)
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "osufusion_b200.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    r = subprocess.run(["gcc", "-std=c99", "-I", str(root / "include"), str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]       # also fails when a ctypes field name does not exist in the C struct
    got = {}
    for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines():
        s, f, v = line.split()
        got[(s, f)] = int(v)
    for cname, cls in pairs.items():
        assert got[(cname, "size")] == C.sizeof(cls), (cname, got[(cname, "size")], C.sizeof(cls))
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)
        # same number of fields: count the declarators of the C struct body
        head = header[:header.index("} " + cname + ";")]
        body = head[head.rindex("typedef struct {") + len("typedef struct {"):]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        n_c = sum(len(stmt.split(",")) for stmt in body.split(";") if stmt.strip())
        assert n_c == len(cls._fields_), (cname, n_c, len(cls._fields_))
