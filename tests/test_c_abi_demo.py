"""The boundary from plain C: examples/c_abi_demo.c builds with gcc against include/wifi_b200.h alone and runs the
IRS_tranceiver-style loopback (mac -> TX -> channel -> RX) through the C ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gnuradio-wifi-imagetransfer_b200")


def _build(tmp_path):
    import importlib
    importlib.import_module("wifi_b200").build.build()
    exe = str(tmp_path / "c_abi_demo")
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c_abi_demo.c"), "-o", exe, "-L", PKG, "-lwifi_b200", "-Wl,-rpath," + PKG, "-lm"])
    return exe


def test_c_demo_builds_and_refuses_to_run_without_a_gpu(tmp_path):
    exe = _build(tmp_path)
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3 and "no CPU path" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("enc,snr", [(3, 27.0), (7, 40.0), (0, 20.0)])
def test_c_demo_loopback(tmp_path, enc, snr):
    exe = _build(tmp_path)
    r = subprocess.run([exe, "24", str(enc), str(snr)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "24 payloads bit-exact" in r.stdout
