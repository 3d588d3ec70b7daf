"""The C++ drop-in front end (include/superbblas.h) exercised by a caller written against the
reference's API (tests/cxx/dropin_test.cpp), run on the GPU."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_cxx_dropin_program():
    exe = os.path.join(HERE, "cxx", "dropin_test")
    if not os.path.exists(exe):
        r = subprocess.run(["make", "-C", os.path.join(HERE, "cxx")], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Everything went ok!" in r.stdout
