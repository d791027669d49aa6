"""Builds the C-side boundary artefacts for the tests (gcc only, no JDK):
  libcvxb_jni.so   jni/cvxb_jni.c against tests/c/jni_stub/jni.h (the subset of the JNI the shim uses)
  drive_abi        tests/c/drive_abi.c: plain C against include/cvxb.h + libcvxb.so
  drive_jni        tests/c/drive_jni.c + fake_jvm.c + the shim: executes every native method through a fake JNIEnv
Outputs go to tests/c/build/ (git-ignored)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CDIR = os.path.join(ROOT, "tests", "c")
OUT = os.path.join(CDIR, "build")
LIBDIR = os.path.join(ROOT, "cvx_b200", "lib")
WARN = ["-std=c11", "-Wall", "-Wextra", "-Werror", "-O1"]


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, "command failed: %s\n%s%s" % (" ".join(cmd), r.stdout, r.stderr)


def build_all():
    os.makedirs(OUT, exist_ok=True)
    inc = ["-I" + os.path.join(ROOT, "include")]
    jni = ["-I" + os.path.join(CDIR, "jni_stub"), "-I" + CDIR]
    link = ["-L" + LIBDIR, "-lcvxb", "-lm", "-Wl,-rpath," + LIBDIR]
    shim = os.path.join(ROOT, "jni", "cvxb_jni.c")
    outs = {"jni_so": os.path.join(OUT, "libcvxb_jni.so"), "drive_abi": os.path.join(OUT, "drive_abi"),
            "drive_jni": os.path.join(OUT, "drive_jni")}
    _run(["gcc"] + WARN + ["-shared", "-fPIC"] + jni + inc + [shim, "-o", outs["jni_so"]] + link)
    _run(["gcc"] + WARN + inc + [os.path.join(CDIR, "drive_abi.c"), "-o", outs["drive_abi"]] + link)
    _run(["gcc"] + WARN + jni + inc + [os.path.join(CDIR, "drive_jni.c"), os.path.join(CDIR, "fake_jvm.c"), shim, "-o",
                                       outs["drive_jni"]] + link)
    return outs
