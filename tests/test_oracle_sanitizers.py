"""The checker itself under AddressSanitizer + UndefinedBehaviorSanitizer (SURVEY.md
section 5: the reference has no sanitizer runs; the new build's oracle gets them).
oracle/sanitize_check.c compiles both oracle C files as one TU with
-fsanitize=address,undefined and drives them over exact-size heap buffers, including a
byte-flip / truncation fuzz of the LZ4 block decoder, the container walk and the Zstd frame
decoder (frames from the real libzstd).  CPU only;
the CUDA side has its own compute-sanitizer run (tools/sanitize.sh, profiles/)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_is_clean_under_asan_and_ubsan(tmp_path):
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    exe = tmp_path / "sanitize_check"
    from oracle import oracle as O
    zstd = ["-DHAVE_LIBZSTD", "-l:libzstd.so.1"] if O.libzstd() is not None else []  # frames for the Zstd fuzz
    build = subprocess.run(
        [gcc, "-std=c11", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
         "-fno-omit-frame-pointer", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "oracle"),
         os.path.join(ROOT, "oracle", "sanitize_check.c"), "-o", str(exe)] + zstd,
        capture_output=True, text=True, timeout=300)
    if build.returncode != 0 and "sanitize" in build.stderr and "cannot find" in build.stderr:
        pytest.skip("toolchain has no sanitizer runtime")
    assert build.returncode == 0, build.stderr
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600, env=env)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "sanitize_check ok" in run.stdout and "ERROR" not in run.stderr
