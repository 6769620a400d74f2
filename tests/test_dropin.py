"""The B200 backends inside the unmodified reference pipeline (oracle/_ref/dropin_test, built by `make -C oracle dropin`
against the reference's own longtail.h and liblongtail_ref.a)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "dropin_test")


def _need_binary():
    if not os.path.exists(BIN):
        if os.path.isdir("/root/reference"):
            from longtail_b200 import build
            build.build()
            subprocess.run(["make", "-s", "-j8", "-C", os.path.join(ROOT, "oracle"), "dropin"], check=True)
        else:
            pytest.skip("oracle/_ref/dropin_test not built (needs /root/reference)")


def test_abi_layout_matches_reference_header():
    """include/longtail_abi.h == reference src/longtail.h, member by member (sizeof + offsetof)"""
    _need_binary()
    r = subprocess.run([BIN, "--abi-only"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "identical layout" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("target", [16, 4096])
def test_b200_backends_inside_reference_pipeline(target):
    """reference Longtail_CreateVersionIndex driving the B200 ChunkerAPI/HashAPI objects, and
    Longtail_B200_CreateVersionIndex, both byte-identical to the all-reference run; golden chunker vector through the API"""
    _need_binary()
    r = subprocess.run([BIN, os.path.join(ROOT, "tests", "golden", "chunker.input"), str(target)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "DROPIN OK" in r.stdout
