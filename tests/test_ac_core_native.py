"""Host unit test of cbc_b200/csrc/ac_core.h (closed-form renormalisation and exact reciprocal division)
against a bit-serial restatement of the reference's coder loops: tests/native/test_ac_core.cpp."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ac_core_closed_form_equals_bit_serial():
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "t")
        subprocess.run(["g++", "-O2", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "cbc_b200", "csrc"),
                        os.path.join(ROOT, "tests", "native", "test_ac_core.cpp"), "-o", exe], check=True)
        p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
        assert p.returncode == 0 and "ac_core ok" in p.stdout, p.stdout + p.stderr
