"""The product's blob parser takes untrusted bytes (rnnoise_model_from_buffer / _from_file): a seeded mutation
fuzzer built with AddressSanitizer + UBSan must run clean (tests/fuzz/fuzz_blob.c)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("model", ["tiny", "default"])
def test_blob_parser_survives_mutations_under_asan(tmp_path, model):
    exe = str(tmp_path / "fuzz_blob")
    cc = subprocess.run(["gcc", "-g", "-O1", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                         "-I", os.path.join(ROOT, "rnnoise_b200", "csrc"), os.path.join(ROOT, "tests", "fuzz", "fuzz_blob.c"),
                         os.path.join(ROOT, "rnnoise_b200", "csrc", "model_blob.c"), "-o", exe], capture_output=True, text=True)
    if cc.returncode != 0:
        pytest.skip("no sanitizer runtime in this toolchain: " + cc.stderr[-200:])
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "models", model + ".bin"), "3000"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout[-300:], r.stderr[-1500:])
    acc, rej = (int(x) for x in r.stdout.split()[1::2])
    assert acc > 0 and rej > 0 and acc + rej == 3000
