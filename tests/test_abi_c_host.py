"""The plain-C host of the C ABI (tests/abi_c/abi_c_check.c): the executable stand-in for the Julia `@ccall` glue.
CPU: it compiles and links against include/b200newton.h + the in-tree library with gcc and refuses to run without a
device (exit code 3: the product has no CPU path).  GPU: config 1 and the ensemble gather through the C-ABI collective."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "abi_c", "abi_c_check")


def _build():
    if not os.path.exists(EXE):
        subprocess.run(["make", "-C", os.path.dirname(EXE), "abi_c_check"], check=True)


def test_c_host_builds_and_refuses_without_device(nls):
    _build()
    if nls.device_count() > 0:
        pytest.skip("a device is present")
    r = subprocess.run([EXE], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_c_host_config1_and_collective():
    _build()
    r = subprocess.run([EXE], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "ABI_C_OK" in r.stdout, r.stdout
    assert "gathered through b200_ens_allgather" in r.stdout
