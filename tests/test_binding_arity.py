"""Every ctypes signature in _lib.py has exactly the arity of the C prototype in include/ctan_sm100.h
(a mismatch is a host-side crash, so it is checked on the CPU)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_signature_arity_matches_header():
    import ctan_b200  # noqa: F401
    from ctan_b200 import _lib
    txt = open(os.path.join(ROOT, "include", "ctan_sm100.h")).read()
    protos = dict(re.findall(r"^int (ctan_\w+)\(([^;]*)\);", txt, re.M))
    for name, args in _lib.SIGNATURES.items():
        c_args = [a for a in protos[name].split(",") if a.strip()]
        assert len(c_args) == len(args) + 1, (name, len(c_args), len(args) + 1)      # +1: the trailing cudaStream_t
        assert "cudaStream_t" in c_args[-1], name
        for a, t in zip(c_args, args):
            is_ptr = "*" in a
            assert is_ptr == (t is _lib.P), (name, a.strip(), t)
