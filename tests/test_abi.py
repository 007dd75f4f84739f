"""The C-ABI library loads on a CPU-only box and exports every entry point include/cbcg.h declares
(no compute calls here: there is no GPU and no CPU fallback to make them on)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "cbcg.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(cbcg_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from cbc_b200 import codec
    lib = codec.load_library()
    names = _declared()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(codec.EXPORTS) == names
    assert lib.cbcg_abi_version() == 1


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    from cbc_b200 import codec
    lib = codec.load_library()
    h = C.c_void_p()
    assert lib.cbcg_create(0, C.byref(h)) == -3          # CBCG_ERR_NO_DEVICE: the product path refuses to run
    assert lib.cbcg_decoded_size(b"x" * 64, 64, None, None) == -8


def test_product_code_never_touches_the_oracle():
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "cbc_b200")):
        if "_build" in d:
            continue
        for fn in files:
            if fn.endswith((".py", ".c", ".h", ".cu", ".cuh", ".cpp")):
                with open(os.path.join(d, fn), errors="replace") as f:
                    t = f.read()
                if re.search(r"oracle_lib|cbc_oracle|cbco_|oracle/", t):
                    bad.append(os.path.join(d, fn))
    assert not bad, bad


def test_ctypes_mirrors_match_the_header_layout(tmp_path):
    """The ctypes structures of cbc_b200/codec.py and batch.py against include/cbcg.h as gcc lays it out: sizes and the
    offset of every field (a header edit that the binding does not follow shows up here, not as a garbled stat)."""
    import subprocess
    import numpy as np
    from cbc_b200 import codec
    from cbc_b200.batch import CBatch
    mirrors = {"cbcg_stats": codec.Stats, "cbcg_encode_opts": codec.EncodeOpts, "cbcg_batch": CBatch}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "cbcg.h"', 'int main(void) {']
    for cname, cls in mirrors.items():
        lines.append(f'    printf("{cname} size %zu\\n", sizeof({cname}));')
        for f, _ in cls._fields_:
            lines.append(f'    printf("{cname} {f} %zu\\n", offsetof({cname}, {f}));')
    lines.append('    printf("cbcg_read_rec size %zu\\n", sizeof(cbcg_read_rec));')
    for f in codec.REC_DTYPE.names:
        lines.append(f'    printf("cbcg_read_rec {f} %zu\\n", offsetof(cbcg_read_rec, {f}));')
    lines += ['    return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = {}
    for ln in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n"):
        if ln:
            s, f, v = ln.split()
            got[(s, f)] = int(v)
    for cname, cls in mirrors.items():
        assert got[(cname, "size")] == C.sizeof(cls), cname
        for f, _ in cls._fields_:
            assert got[(cname, f)] == getattr(cls, f).offset, (cname, f)
    assert got[("cbcg_read_rec", "size")] == codec.REC_DTYPE.itemsize == 16
    for f in codec.REC_DTYPE.names:
        assert got[("cbcg_read_rec", f)] == codec.REC_DTYPE.fields[f][1], f
    assert np.dtype(codec.SYM_DTYPE).itemsize == 8
