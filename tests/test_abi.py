"""The C-ABI library loads and exports every symbol include/cube_gpu.h declares (no compute calls without a GPU),
and the product path fails loudly instead of falling back to the CPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "cube_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cube_gpu_\w+)\s*\(", src)))


def test_header_declares_the_step_calls():
    syms = header_symbols()
    for s in ("cube_gpu_init", "cube_gpu_upload", "cube_gpu_update_x", "cube_gpu_buffer", "cube_gpu_particle_mesh",
              "cube_gpu_download", "cube_gpu_finalize", "cube_gpu_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from cafproject_b200 import cube
    L = cube.load_library()
    for s in header_symbols():
        assert hasattr(L, s), "libcubegpu.so does not export %s" % s
    assert sorted(cube.ABI_SYMBOLS) == header_symbols()


def test_params_struct_layout_matches_header():
    """cube_params is plain int32/float fields: 3+1+1+1+1+1+2+1 ints, 2 floats, 2 ints, 4 reserved = 76 bytes."""
    from cafproject_b200.cube import CubeParams
    assert C.sizeof(CubeParams) == 4 * (3 + 1 + 1 + 1 + 1 + 1 + 2 + 1 + 2 + 2 + 4)


def test_no_cpu_fallback(tables):
    """Without a CUDA device cube_gpu_init must fail with a message, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from cafproject_b200.cube import CubeGPU, CubeGPUError
    fk, ck = tables
    with pytest.raises(CubeGPUError, match="no CUDA device|CUDA"):
        CubeGPU(24, 2, fk, ck, tanf_lut=np.zeros(65536, np.float32))


def test_product_does_not_import_the_oracle():
    """Only tests/, smoke() and bench.py's cpu_baseline leg may touch oracle/."""
    pkg = os.path.join(ROOT, "cafproject_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "cube_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_init_rejects_bad_configuration(tables):
    """Error behaviour mirrors the reference's `stop` messages (particle_initialization.f90:14-18)."""
    import torch
    from cafproject_b200 import cube
    L = cube.load_library()
    p = cube.CubeParams()
    p.nn[:] = (1, 1, 1)
    p.nnt, p.nc, p.ncell, p.ncb, p.izipx, p.izipv, p.np_nc = 2, 24, 4, 6, 4, 2, 2   # izipx = 4 does not exist (universe*.fh: 1 or 2)
    fk, ck = tables
    lut = np.zeros(65536, np.float32)
    h = C.c_void_p()
    rc = L.cube_gpu_init(C.byref(p), fk.ctypes.data, ck.ctypes.data, lut.ctypes.data, None, C.byref(h))
    assert rc != 0 and b"zip format incompatable" in L.cube_gpu_last_error()


def _strip_c_comments(src):
    return re.sub(r"/\*.*?\*/", "", src, flags=re.S)


def test_fortran_module_mirrors_the_header():
    """fortran/cube_gpu.f90 cannot be compiled here (no Fortran compiler), so keep it in step with include/cube_gpu.h
    textually: the bind(C) cube_params has the header's fields in the header's order (same counts), and every
    interface names an exported symbol with the header's number of arguments."""
    hdr = _strip_c_comments(open(os.path.join(ROOT, "include", "cube_gpu.h")).read())
    f90 = open(os.path.join(ROOT, "fortran", "cube_gpu.f90")).read()
    f90 = "\n".join(line.split("!")[0] for line in f90.splitlines())
    # struct fields
    body = re.search(r"typedef struct cube_params \{(.*?)\} cube_params;", hdr, re.S).group(1)
    c_fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        ctype, names = decl.split(None, 1)
        for nm in names.split(","):
            m = re.match(r"\s*(\w+)(?:\[(\d+)\])?", nm)
            c_fields.append((m.group(1), int(m.group(2) or 1), "float" if ctype == "float" else "int"))
    tbody = re.search(r"type, bind\(C\) :: cube_params(.*?)end type", f90, re.S).group(1)
    f_fields = []
    for line in tbody.strip().splitlines():
        m = re.match(r"\s*(integer\(c_int32_t\)|real\(c_float\))\s*::\s*(.*)", line)
        assert m, line
        for nm in m.group(2).split(","):
            mm = re.match(r"\s*(\w+)(?:\((\d+)\))?", nm)
            f_fields.append((mm.group(1), int(mm.group(2) or 1), "float" if m.group(1).startswith("real") else "int"))
    assert f_fields == c_fields
    # interfaces: name and arity
    c_protos = {m.group(1): m.group(2) for m in re.finditer(r"\b(cube_gpu_\w+)\s*\(([^)]*)\)\s*;", hdr)}
    arity = lambda a: 0 if a.strip() in ("", "void") else len(a.split(","))
    found = re.findall(r"function\s+(cube_gpu_\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name=\"(\w+)\"\)", f90)
    assert len(found) >= 12
    for fname, fargs, cname in found:
        assert fname == cname and cname in c_protos, cname
        assert arity(fargs) == arity(c_protos[cname]), cname
    step_calls = {"cube_gpu_init", "cube_gpu_upload", "cube_gpu_update_x", "cube_gpu_buffer", "cube_gpu_particle_mesh",
                  "cube_gpu_download", "cube_gpu_finalize", "cube_gpu_last_error"}
    assert step_calls <= {c for _, _, c in found}
