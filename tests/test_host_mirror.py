"""The C++ host mirror of the generated graph surface compiles, links against the C-ABI library
and (on a GPU) renders audible output from a MIDI note, like the reference's example smoke tests."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    from oscen_b200 import build
    lib = build.build()
    exe = str(tmp_path / "host_mirror_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "csrc", "host_mirror_test.cpp"),
                           lib, f"-Wl,-rpath,{os.path.dirname(lib)}", "-ldl", "-lpthread", "-lrt"])
    return exe


def test_host_mirror_builds_and_refuses_without_gpu(tmp_path):
    import conftest
    exe = _build(tmp_path)
    rc = subprocess.call([exe])
    assert rc == (0 if conftest.HAS_GPU else 77)


@pytest.mark.gpu
def test_host_mirror_midi_note_makes_sound(tmp_path):
    exe = _build(tmp_path)
    assert subprocess.call([exe]) == 0
