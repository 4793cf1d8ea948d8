"""Host code under AddressSanitizer + UndefinedBehaviorSanitizer: the BDD collection (bdd_b200/csrc/host/bdd_collection.hpp, split.hpp), the
.lp-side BDD builder (bdd_solver_native.hpp) and the layout builder (bdd_b200/csrc/layout.hpp) fed with random constraint systems
(tests/cpp/fuzz_*.cpp), the .lp reader with mutated fixture files.  CPU only; skipped where g++ has no sanitizer runtime."""
import glob
import os
import subprocess
import sysconfig

import pytest

from conftest import ROOT

CPP = os.path.join(ROOT, "tests", "cpp")


@pytest.mark.parametrize("name, flags", [("fuzz_collection", []), ("fuzz_layout", ["-fopenmp"]), ("fuzz_lp", [])])
def test_host_code_is_clean_under_asan_and_ubsan(tmp_path, name, flags):
    if not os.path.exists(os.path.join(ROOT, "bdd_b200", "libbdd_b200.so")):
        pytest.skip("libbdd_b200.so not built")
    json_inc = os.path.join(sysconfig.get_paths()["purelib"], "include", "cudnn_frontend", "thirdparty")
    exe = str(tmp_path / name)
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", *flags, "-I" + json_inc,
           "-o", exe, os.path.join(CPP, name + ".cpp"), "-L" + os.path.join(ROOT, "bdd_b200"), "-lbdd_b200", "-Wl,-rpath," + os.path.join(ROOT, "bdd_b200")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    if r.returncode != 0 and ("asan" in r.stdout or "ubsan" in r.stdout or "sanitize" in r.stdout):
        pytest.skip("no sanitizer runtime for g++ here")
    assert r.returncode == 0, r.stdout[-3000:]
    env = dict(os.environ, OMP_NUM_THREADS="4", ASAN_OPTIONS="detect_leaks=1:abort_on_error=0")
    args = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.lp"))) if name == "fuzz_lp" else []
    r = subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    assert r.returncode == 0 and r.stdout.startswith("ok "), r.stdout[-3000:]
