"""The C ABI consumed from plain C (no Python, no torch in the process): the program is what a
Rust `-sys` crate would link against.  On a CPU box it must see GSB_ERR_NO_DEVICE."""
import os
import subprocess

import pytest

import gsearch_b200 as g

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(tmp_path):
    exe = str(tmp_path / "c_consumer")
    libdir = os.path.join(ROOT, "gsearch_b200")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, "-std=c11", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_consumer.c"), "-o", exe, "-L", libdir, "-lgsearch_b200",
                    "-Wl,-rpath," + libdir], check=True)
    return exe


@pytest.mark.skipif(g.device_count() > 0, reason="CPU-box behaviour")
def test_c_program_links_and_sees_no_device(tmp_path):
    r = subprocess.run([build(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 2, r.stdout + r.stderr
    assert "sm_100a" in r.stdout and "no device" in r.stdout


@pytest.mark.gpu
def test_c_program_runs_the_whole_path(tmp_path):
    r = subprocess.run([build(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok: d(0,1)=0" in r.stdout
