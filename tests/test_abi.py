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
