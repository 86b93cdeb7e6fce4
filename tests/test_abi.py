"""CPU tests of the C-ABI boundary: the library loads, exports every symbol the header declares, and fails
loudly (no CPU fallback) when there is no GPU."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "scalaopt_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(so_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(native):
    lib = native.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    missing = [s for s in declared if not hasattr(lib, s)]
    assert missing == []
    assert sorted(native.EXPORTS) == declared
    assert lib.so_abi_version() == 2


def test_no_cpu_fallback_without_a_gpu(native):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(native.NativeError) as e:
        native.Context(1)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure; nothing under scalaopt_b200/ may reference it."""
    pkg = os.path.join(ROOT, "scalaopt_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", ".c")):
                src = open(os.path.join(dirpath, fn), errors="ignore").read()
                if re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M) or "liboracle" in src or "oracle.h" in src:
                    offenders.append(os.path.join(dirpath, fn))
    assert offenders == []


# ---- the reference-side binding under integration/ (Scala + JNI): no JVM or JDK here, so what CAN be checked is checked ----
SCALA_DIR = os.path.join(ROOT, "integration", "scala", "com", "github", "bruneli", "scalaopt", "gpu")
JNI_C = os.path.join(ROOT, "integration", "jni", "scalaopt_jni.c")


def _native_defs():
    text = open(os.path.join(SCALA_DIR, "NativeLib.scala")).read()
    return text, sorted(re.findall(r"@native\s+def\s+(\w+)", text))


def test_jni_glue_type_checks_against_the_c_header():
    """gcc -fsyntax-only of the JNI glue against include/scalaopt_cuda.h, with a stand-in for the JDK's <jni.h>
    (tests/jni_stub/jni.h): every so_* call in the glue has the argument count and types the header declares."""
    import subprocess
    proc = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter",
                           "-I" + os.path.join(ROOT, "tests", "jni_stub"), "-I" + os.path.join(ROOT, "include"), JNI_C],
                          capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    src = re.sub(r"/\*.*?\*/", "", open(JNI_C).read(), flags=re.S)
    called = set(re.findall(r"\b(so_[a-z_0-9]+)\s*\(", src))
    assert called and called <= set(_declared_symbols())
    # blocking calls must not sit inside JNI critical sections (the collector would be held off for their duration)
    assert "PrimitiveArrayCritical(" not in src


def test_scala_native_declarations_match_the_jni_glue():
    text, natives = _native_defs()
    glue = sorted(re.findall(r"JFN\((\w+)\)\s*\(", open(JNI_C).read()))
    assert natives == glue and len(natives) >= 18
    # same number of parameters on both sides (JNI adds JNIEnv* and the receiver), and the same kinds in the same order
    kinds = {"Int": "jint", "Long": "jlong", "Double": "jdouble", "Boolean": "jboolean", "String": "jstring",
             "Array[Double]": "jdoubleArray", "Array[Int]": "jintArray", "java.nio.ByteBuffer": "jobject"}
    csrc = open(JNI_C).read()
    for name, params in re.findall(r"@native\s+def\s+(\w+)\(([^)]*)\)", text):
        want = [kinds[t.split(":")[1].strip()] for t in params.split(",")] if params.strip() else []
        cparams = re.search(r"JFN\(" + name + r"\)\s*\(([^)]*)\)", csrc).group(1).split(",")
        got = [c.strip().rsplit(" ", 1)[0].strip() for c in cparams][2:]
        assert got == want, (name, got, want)
    # and the Scala sources only use what NativeLib declares (@native methods or its constants)
    consts = set(re.findall(r"\bval\s+(\w+)\s*=", text))
    for fn in os.listdir(SCALA_DIR):
        if fn == "NativeLib.scala":
            continue
        used = set(re.findall(r"NativeLib\.(\w+)", open(os.path.join(SCALA_DIR, fn)).read()))
        assert used <= set(natives) | consts, (fn, used - set(natives) - consts)
