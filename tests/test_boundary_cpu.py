"""The drop-in boundary below the Scala side, checked without a GPU and without a JDK:
  * jni/cvxb_jni.c compiles warning-free (-Wall -Wextra -Werror) against the JNI subset it uses and exports exactly
    the native methods scala/cvx/CvxbNative.scala declares;
  * it never opens a JNI critical region (cvxb_* calls block and allocate);
  * a plain C program links against include/cvxb.h + libcvxb.so and fails loudly without a GPU (no CPU path)."""
import os
import re
import subprocess

from tests import boundary_build as bb

ROOT = bb.ROOT


def _scala_natives():
    src = open(os.path.join(ROOT, "scala", "cvx", "CvxbNative.scala")).read()
    return set(re.findall(r"@native\s+def\s+(\w+)", src))


def test_shim_compiles_and_matches_scala_natives():
    outs = bb.build_all()
    nm = subprocess.run(["nm", "-D", "--defined-only", outs["jni_so"]], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"Java_cvx_CvxbNative_(\w+)", nm))
    natives = _scala_natives()
    assert natives, "no @native methods found in CvxbNative.scala"
    assert exported == natives, "shim exports %s, Scala declares %s" % (sorted(exported - natives), sorted(natives - exported))


def test_shim_has_no_critical_regions_and_builds_exceptions_properly():
    src = open(os.path.join(ROOT, "jni", "cvxb_jni.c")).read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    assert "GetPrimitiveArrayCritical" not in code
    # LinSolveException has no (String) constructor (LinSolveException.scala:11-17): NewObject with the 4-argument one
    assert "Lbreeze/linalg/DenseMatrix;Lbreeze/linalg/DenseVector;Lbreeze/linalg/DenseMatrix;Ljava/lang/String;)V" in code
    assert re.search(r'ThrowNew\([^;]*LinSolveException', code) is None
    # every FindClass result is checked before use
    for m in re.finditer(r"(\w+)\s*=\s*\(\*env\)->FindClass", code):
        var = m.group(1)
        tail = code[m.end():m.end() + 400]
        assert re.search(r"!\s*%s\b|%s\s*\?|if\s*\(\s*%s\b" % (var, var, var), tail), "FindClass result %s used unchecked" % var


def test_scala_natives_match_jni_arity():
    """Argument counts of the Scala @native declarations equal those of the C functions (+ env, class)."""
    scala = open(os.path.join(ROOT, "scala", "cvx", "CvxbNative.scala")).read()
    csrc = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "jni", "cvxb_jni.c")).read(), flags=re.S)
    for name, args in re.findall(r"@native\s+def\s+(\w+)\s*\(([^)]*)\)", scala, flags=re.S):
        n_scala = len([a for a in args.split(",") if a.strip()])
        m = re.search(r"Java_cvx_CvxbNative_%s\s*\(([^)]*)\)" % name, csrc, flags=re.S)
        assert m, name
        n_c = len([a for a in m.group(1).split(",") if a.strip()]) - 2
        assert n_scala == n_c, "%s: Scala declares %d arguments, the shim takes %d" % (name, n_scala, n_c)


def test_c_driver_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    outs = bb.build_all()
    r = subprocess.run([outs["drive_abi"]], capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr and "no CPU path" in r.stderr, r.stderr
