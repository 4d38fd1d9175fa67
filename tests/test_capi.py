"""The C-ABI library loads, exports every symbol include/lvpp_b200.h declares, and its structs have
the layout the ctypes binding assumes.  No compute calls (no GPU here)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

from proximalgalerkin_b200 import _capi

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "lvpp_b200.h"


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(lvpp_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _capi.SIGNATURES, f"{n} has no ctypes prototype"
    assert sorted(_capi.SIGNATURES) == names
    assert lib.lvpp_version() == 100


def test_struct_layout_matches_c(tmp_path, lib):
    src = tmp_path / "layout.c"
    fields = {"lvpp_obstacle_desc": [f for f, _ in _capi.ObstacleDesc._fields_],
              "lvpp_newton_opts": [f for f, _ in _capi.NewtonOpts._fields_],
              "lvpp_stats": [f for f, _ in _capi.Stats._fields_]}
    body = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for st, fl in fields.items():
        body.append(f'printf("{st} %zu\\n", sizeof({st}));')
        for f in fl:
            body.append(f'printf("{st}.{f} %zu\\n", offsetof({st}, {f}));')
    body.append("return 0;}")
    src.write_text("\n".join(body))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", str(src), "-o", str(exe)], check=True)
    out = dict(line.rsplit(" ", 1) for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for st, cls in (("lvpp_obstacle_desc", _capi.ObstacleDesc), ("lvpp_newton_opts", _capi.NewtonOpts), ("lvpp_stats", _capi.Stats)):
        assert int(out[st]) == C.sizeof(cls)
        for f, _ in cls._fields_:
            assert int(out[f"{st}.{f}"]) == getattr(cls, f).offset, (st, f)


def test_argument_validation_without_gpu(lib):
    assert lib.lvpp_set_alpha(None, 1.0) == _capi.E_INVALID
    assert b"null handle" in lib.lvpp_last_error()
    assert lib.lvpp_destroy(None) == _capi.OK
    assert lib.lvpp_get_stats(None, None) == _capi.E_INVALID
    import torch

    if not torch.cuda.is_available():
        assert lib.lvpp_device_count() == _capi.E_NOGPU
        d, h = _capi.ObstacleDesc(), _capi.H()
        assert lib.lvpp_create(C.byref(d), C.byref(h)) == _capi.E_NOGPU
