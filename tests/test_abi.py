"""C-ABI surface: the library loads without a GPU, exports every symbol include/ptb.h declares, and its structs
have the reference's layout (optixSphere.h), pinned by tests/golden/abi_layout.json (written by oracle/_ref/ref_probe,
i.e. the reference header compiled with the real CUDA vector types)."""
import ctypes as C
import json
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLD = json.loads((ROOT / "tests" / "golden" / "abi_layout.json").read_text())


def test_exports_match_header(ptb):
    header = (ROOT / "include" / "ptb.h").read_text()
    declared = set(re.findall(r"\b(ptb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(ptb.EXPORTS), declared ^ set(ptb.EXPORTS)
    L = ptb.lib()
    for name in declared:
        assert hasattr(L, name), f"libptb.so does not export {name}"
    assert b"sm_100a" in L.ptb_version()


def test_struct_layout_matches_reference(ptb):
    def offs(cls):
        return {n: getattr(cls, n).offset for n, _ in cls._fields_ if not n.startswith("_")}
    assert C.sizeof(ptb.Params) == GOLD["sizeof.Params"] == 120
    for f, o in offs(ptb.Params).items():
        assert o == GOLD[f"Params.{f}"], f
    assert C.sizeof(ptb.HitGroupData) == GOLD["sizeof.HitGroupData"] == 168
    for f, o in offs(ptb.HitGroupData).items():
        assert o == GOLD[f"HitGroupData.{f}"], f
    assert C.sizeof(ptb.TriangleData) == GOLD["sizeof.TriangleData"] == 128
    for f, o in offs(ptb.TriangleData).items():
        assert o == GOLD[f"TriangleData.{f}"], f


def test_header_structs_compile_to_reference_layout(tmp_path):
    """The C header itself (not just the ctypes mirror): compile a probe with gcc (as C) and g++ and compare."""
    src = tmp_path / "probe.c"
    fields = [k for k in GOLD if "." in k and not k.startswith(("sizeof", "alignof"))]
    keep = [k for k in fields if k.split(".")[0] in ("TriangleData", "Params", "MissData", "HitGroupData")]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "ptb.h"', "int main(void){"]
    for k in keep:
        s, f = k.split(".")
        lines.append(f'printf("{k} %zu\\n", offsetof(ptb_{s}, {f}));')
    for s in ("TriangleData", "Params", "MissData", "HitGroupData"):
        lines.append(f'printf("sizeof.{s} %zu\\n", sizeof(ptb_{s}));')
    lines.append("return 0;}")
    src.write_text("\n".join(lines))
    for cc, std in (("/usr/bin/gcc", "-std=c11"), ("/usr/bin/g++", "-std=c++17")):
        exe = tmp_path / ("probe_" + Path(cc).name)
        args = [cc, std, "-I", str(ROOT / "include"), "-o", str(exe)] + (["-x", "c++"] if "g++" in cc else []) + [str(src)]
        subprocess.run(args, check=True)
        out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
        for k, v in out.items():
            assert int(v) == GOLD[k], (cc, k, v, GOLD[k])


def test_ctypes_mirrors_match_header(ptb, tmp_path):
    """Every configuration / statistics struct of include/ptb.h against its ctypes mirror: size and every field offset
    (a field appended to the header but not to the binding would otherwise be read as garbage)."""
    pairs = {"ptb_render_cfg": ptb.RenderCfg, "ptb_build_cfg": ptb.BuildCfg, "ptb_build_stats": ptb.BuildStats,
             "ptb_launch_stats": ptb.LaunchStats, "ptb_material_info": ptb.MaterialInfo}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "ptb.h"', "int main(void){"]
    for cname, cls in pairs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for f, _ in cls._fields_:
            lines.append(f'printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines.append("return 0;}")
    src = tmp_path / "cfg_probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "cfg_probe"
    subprocess.run(["/usr/bin/gcc", "-std=c11", "-I", str(ROOT / "include"), "-o", str(exe), str(src)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in pairs.items():
        assert int(out[cname]) == C.sizeof(cls), cname
        for f, _ in cls._fields_:
            assert int(out[f"{cname}.{f}"]) == getattr(cls, f).offset, (cname, f)


def test_golden_layout_matches_live_reference_probe(oh):
    if not oh.REF_PROBE.exists():
        pytest.skip("oracle/_ref/ref_probe not built (no /root/reference on this box)")
    live = json.loads(subprocess.run([str(oh.REF_PROBE), "layout"], capture_output=True, text=True, check=True).stdout)
    live.pop("end")
    assert live == GOLD


def test_no_device_fails_loudly(ptb):
    """The product has no CPU fallback: without a usable CUDA device the context cannot be created."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ptb.PtbError) as e:
        ptb.Context(0)
    assert e.value.code == ptb.PTB_ERR_NO_DEVICE
    assert "no CPU path" in str(e.value)


def test_product_does_not_touch_the_oracle():
    """Nothing in the package or include/ may include, import, load or execute anything under oracle/
    (the oracle is test infrastructure).  Comments and docstrings may cite it."""
    import ast
    pkg = ROOT / "szakdolgozat_pathtracer_b200"
    files = [f for f in list(pkg.rglob("*")) + list((ROOT / "include").rglob("*")) if f.is_file()]
    assert any(f.suffix == ".cu" for f in files)
    for f in files:
        if f.suffix in (".h", ".cuh", ".cu", ".cpp"):
            txt = f.read_text(errors="replace")
            assert not re.search(r'#\s*include\s*[<"][^>"]*(oracle|orchelp|ref_shim)', txt), f
            code = re.sub(r"/\*.*?\*/", "", re.sub(r"//[^\n]*", "", txt), flags=re.S)
            assert "oracle" not in code and "orc_" not in code, f
        elif f.name == "Makefile":
            code = re.sub(r"#[^\n]*", "", f.read_text())
            assert "oracle" not in code, f
        elif f.suffix == ".py":
            tree = ast.parse(f.read_text())
            doc_nodes = set()
            for node in ast.walk(tree):
                if isinstance(node, (ast.Module, ast.ClassDef, ast.FunctionDef)) and node.body and isinstance(node.body[0], ast.Expr) \
                        and isinstance(node.body[0].value, ast.Constant) and isinstance(node.body[0].value.value, str):
                    doc_nodes.add(id(node.body[0].value))
            for node in ast.walk(tree):
                if isinstance(node, ast.Constant) and isinstance(node.value, str) and id(node) not in doc_nodes:
                    assert "oracle" not in node.value and "orchelp" not in node.value, (f, node.value)
                if isinstance(node, (ast.Import, ast.ImportFrom)):
                    names = [a.name for a in node.names] + [getattr(node, "module", "") or ""]
                    assert not any("oracle" in n or "orchelp" in n for n in names), f


def test_accum_raw_round_trip(ptb, tmp_path):
    """ptb_save_accum_raw / ptb_load_accum_raw: bit-exact round trip incl. NaN payloads and the header."""
    rng = np.random.default_rng(3)
    a = rng.standard_normal((7, 13, 4)).astype(np.float32)
    a.view(np.uint32)[0, 0, 0] = 0x7FC12345
    f = tmp_path / "accum.ptba"
    ptb.save_accum_raw(f, a)
    raw = f.read_bytes()
    assert raw[:4] == b"PTBA" and np.frombuffer(raw[4:16], np.uint32).tolist() == [1, 13, 7] and len(raw) == 16 + a.nbytes
    b = ptb.load_accum_raw(f)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    f.write_bytes(raw[:40])
    with pytest.raises(ptb.PtbError):
        ptb.load_accum_raw(f)
    with pytest.raises(ptb.PtbError):
        ptb.load_accum_raw(tmp_path / "missing.ptba")
