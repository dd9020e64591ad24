"""The Java side of the boundary ships as source (java/): there is no JDK in this image, so the CPU
suite checks what it can — the JNI glue compiles against a minimal jni.h, every native method of
cc.mrlda.b200.MrLdaB200 has its C function, and the glue only calls symbols the ABI header declares."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JAVA = os.path.join(ROOT, "java", "cc", "mrlda", "b200")
GLUE = os.path.join(ROOT, "java", "jni", "mrlda_b200_jni.c")
HEADER = os.path.join(ROOT, "include", "mrlda_b200.h")


def test_jni_glue_compiles_against_stub_jni_h():
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-Wall", "-Wextra", "-Werror",
                        "-I", os.path.join(ROOT, "tests", "jni_stub"), "-I", os.path.join(ROOT, "include"), GLUE],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_every_native_method_has_a_jni_function():
    src = open(os.path.join(JAVA, "MrLdaB200.java")).read()
    natives = re.findall(r"\bnative\s+[\w\[\]]+\s+(\w+)\s*\(", src)
    assert len(natives) >= 14
    glue = open(GLUE).read()
    defined = set(re.findall(r"Java_cc_mrlda_b200_MrLdaB200_(\w+)\s*\(", glue))
    assert set(natives) == defined, (sorted(set(natives) - defined), sorted(defined - set(natives)))


def test_glue_calls_only_declared_abi_symbols():
    header = open(HEADER).read()
    declared = set(re.findall(r"\b(mrlda_\w+)\s*\(", header))
    used = set(re.findall(r"\b(mrlda_\w+)\s*\(", open(GLUE).read()))
    assert used and used <= declared, sorted(used - declared)


def test_helper_classes_and_patch_are_present():
    for name in ("MrLdaB200.java", "CsrCorpus.java", "BetaFile.java", "GammaDir.java"):
        text = open(os.path.join(JAVA, name)).read()
        assert "package cc.mrlda.b200;" in text and text.count("{") == text.count("}")
    patch = open(os.path.join(ROOT, "java", "VariationalInference.patch")).read()
    assert patch.startswith("--- a/src/main/java/cc/mrlda/VariationalInference.java")
    # hunk headers are consistent with their bodies
    for m in re.finditer(r"^@@ -(\d+),(\d+) \+(\d+),(\d+) @@.*\n((?:[ +\-].*\n|\n)*)", patch, re.M):
        body = m.group(5).splitlines()
        old = sum(1 for l in body if l[:1] in (" ", "-") or l == "")
        new = sum(1 for l in body if l[:1] in (" ", "+") or l == "")
        assert (old, new) == (int(m.group(2)), int(m.group(4))), m.group(0)[:80]


def test_patch_applies_to_the_reference_driver(tmp_path):
    """Where the reference tree is present (this container, not the GPU box) the patch must apply
    to the unmodified cc.mrlda.VariationalInference with no fuzz and no rejected hunk."""
    import shutil

    import pytest
    ref = "/root/reference/src/main/java/cc/mrlda/VariationalInference.java"
    if not os.path.exists(ref) or shutil.which("patch") is None:
        pytest.skip("reference tree or patch(1) not available")
    dst = tmp_path / "src" / "main" / "java" / "cc" / "mrlda"
    dst.mkdir(parents=True)
    shutil.copy(ref, dst / "VariationalInference.java")
    r = subprocess.run(["patch", "-p1", "--fuzz=0", "-i", os.path.join(ROOT, "java", "VariationalInference.patch")],
                       cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0 and "FAILED" not in r.stdout and "fuzz" not in r.stdout, r.stdout + r.stderr
    patched = (dst / "VariationalInference.java").read_text()
    original = open(ref).read()  # (its own brace count is off by one: a brace in a string literal)
    assert "MrLdaB200" in patched
    assert patched.count("{") - patched.count("}") == original.count("{") - original.count("}")
    assert not list(tmp_path.rglob("*.rej"))


def _glue_functions():
    glue = open(GLUE).read()
    parts = re.split(r"^JNIEXPORT ", glue, flags=re.M)[1:]
    return {re.search(r"MrLdaB200_(\w+)\s*\(", p).group(1): p for p in parts}


def test_glue_checks_array_lengths_before_the_abi_call():
    """The C ABI trusts pointer extents; a short Java array must become an IOException in the glue."""
    fns = _glue_functions()
    for name in ("setCorpus", "setAlpha", "setBeta", "getBeta", "getGamma"):
        body = fns[name]
        assert "bad_length(" in body, name
        call = re.search(r"\bmrlda_(?:set|get)_\w+\(ctx", body)
        assert call and body.index("bad_length(") < call.start(), name  # checked before the ABI call
    assert "rowPtr[D]" in fns["setCorpus"]  # termId / count against Z


def test_glue_calls_no_jni_function_inside_a_critical_region():
    """Between GetPrimitiveArrayCritical and the last Release no other JNI call is allowed (JNI
    specification); throw_io / check / set_docs call JNI and must come after the last unpin."""
    for name, body in _glue_functions().items():
        if "pin(env" not in body:
            continue
        first = body.index("pin(env")
        last = body.rindex("unpin(env")
        region = body[first:last]
        for banned in ("throw_io(", "check(env", "set_docs(", "GetArrayLength", "_of(env"):
            # a pin that fails on the FIRST array holds nothing yet; that early exit is the only allowed one
            hits = [m.start() for m in re.finditer(re.escape(banned), region)]
            for h in hits:
                before = region[:h]
                assert before.count("pin(env") - before.count("unpin(env") == 1 and "kPinFailed" in region[h:h + 40], \
                    (name, banned)
