"""The C ABI and the JNI shim driven from C on the GPU (VERDICT r1 #7 / ADVICE r1): a plain C program against
include/cvxb.h, and every native method of jni/cvxb_jni.c executed through a fake JNIEnv (tests/c/fake_jvm.c) --
including which exception class and constructor the shim uses on the failure paths."""
import subprocess

import pytest

from tests import boundary_build as bb

pytestmark = pytest.mark.gpu


def test_c_program_drives_the_abi():
    outs = bb.build_all()
    r = subprocess.run([outs["drive_abi"]], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "drive_abi ok" in r.stdout, r.stdout + r.stderr


def test_jni_shim_executes_through_fake_jvm():
    outs = bb.build_all()
    r = subprocess.run([outs["drive_jni"]], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "drive_jni ok" in r.stdout, r.stdout + r.stderr
    assert "LinSolveException via (Lbreeze/linalg/DenseMatrix;" in r.stdout
