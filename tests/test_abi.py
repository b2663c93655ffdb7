"""The C-ABI libraries load and export every symbol the headers declare; without a GPU the product fails loudly."""
import ctypes as C
import re
import subprocess

import pytest

import helpers as H

capi = H.capi


def header_functions(name):
    text = (H.ROOT / "include" / name).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pe[h]?_[a-z0-9_]+)\s*\(", text)))


def test_device_library_exports_every_declared_symbol():
    lib = capi.load_device()
    declared = header_functions("poroel.h")
    assert set(declared) == set(capi.DEVICE_SYMBOLS)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.pe_version() == 100


def test_host_library_exports_every_declared_symbol():
    lib = capi.load_host()
    declared = [s for s in header_functions("poroel_host.h") if s.startswith("peh_") and not s.endswith(("_view", "_report"))]
    assert set(declared) == set(capi.HOST_SYMBOLS)
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_device_library_is_sm100a_only_and_has_no_oracle_dependency():
    lib = H.ROOT / "poroelasticity-dealii_b200" / "lib" / "libporoel.so"
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", str(lib)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
    needed = subprocess.run(["readelf", "-d", str(lib)], capture_output=True, text=True).stdout
    assert "oracle" not in needed
    host = subprocess.run(["readelf", "-d", str(lib.with_name("libporoel_host.so"))], capture_output=True, text=True).stdout
    assert "libporoel.so" in host and "oracle" not in host


def test_struct_layouts_match_the_header():
    # sizes computed from the C declarations (8-byte aligned PODs)
    assert C.sizeof(capi.PeParams) == 8 * 4 + 12 * 8
    assert C.sizeof(capi.PeStats) == 14 * 8 + 6 * 8 + 4 * 8 + 4 * 8 + 8 + 6 * 8 + 4 * 8 + 10 * 8


def test_product_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.BackendError) as e:
        capi.create_device_backend(0)
    assert e.value.status == capi.PE_ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_ctypes_struct_offsets_match_a_c_compile_of_the_header(tmp_path):
    """Compile a C program against include/poroel.h and include/poroel_host.h and compare every field offset with the ctypes mirror."""
    def emit(struct, cname, fields):
        lines = [f'  printf("{struct} size %zu\\n", sizeof({cname}));']
        for f in fields:
            lines.append(f'  printf("{struct} {f} %zu\\n", offsetof({cname}, {f}));')
        return "\n".join(lines)

    mirrors = {"pe_params": capi.PeParams, "pe_stats": capi.PeStats, "peh_step_report": capi.StepReport, "peh_mesh_view": capi.MeshView,
               "peh_dofs_view": capi.DofsView, "peh_input_view": capi.InputView, "peh_part_field_view": capi.PartFieldView, "peh_part_view": capi.PartView,
               "peh_constraints_view": capi.ConstraintsView}
    body = "\n".join(emit(name, name, [f for f, _ in cls._fields_]) for name, cls in mirrors.items())
    src = tmp_path / "offsets.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "poroel_host.h"\nint main(void) {\n' + body + "\n  return 0;\n}\n")
    exe = tmp_path / "offsets"
    subprocess.check_call(["gcc", "-std=c11", "-I", str(H.ROOT / "include"), str(src), "-o", str(exe)])  # the headers are plain C
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    seen = 0
    for line in out.splitlines():
        struct, field, value = line.split()
        cls = mirrors[struct]
        if field == "size":
            assert C.sizeof(cls) == int(value), (struct, C.sizeof(cls), value)
        else:
            assert getattr(cls, field).offset == int(value), (struct, field)
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in mirrors.values())
