"""GPU: a plain-C program against the C ABI (no Python / torch in the call path)."""
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_plain_c_client(tmp_path):
    exe = str(tmp_path / "c_abi_client")
    libdir = os.path.join(ROOT, "aqml_b200", "lib")
    subprocess.check_call(["/usr/bin/gcc", "-O1", "-o", exe, os.path.join(ROOT, "tests", "c_abi_client.c"),
                           "-L" + libdir, "-laqml_b200", "-lm", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True)
    print(out.stdout, out.stderr)
    assert out.returncode == 0, out.stdout + out.stderr
