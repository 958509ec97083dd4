"""AddressSanitizer + UBSan over the host-only code of libimpgx (partition bookkeeping, .impg files, the BEDPE / PAF
merge with its CIGAR surgery, the text parsers): the sources are compiled as plain C++ with
-fsanitize=address,undefined and driven by tests/sanitize_host.cpp / sanitize_parsers.cpp with random partition runs
(termination, every base partitioned exactly once), damaged index files, random alignment chains and mutated text."""
import glob
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "impg_b200", "csrc")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")


@pytest.mark.skipif(CXX is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")),
                    reason="needs g++ and the CUDA headers")
def test_host_code_is_clean_under_asan_and_ubsan(tmp_path):
    flags = [CXX, "-std=c++17", "-g", "-O1", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
             "-fno-omit-frame-pointer", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-I" + CUDA_INC]
    objs = []
    for src in ("partition.cu", "impg_file.cu", "host_output.cu", "index_host.cu", "api.cu"):
        obj = str(tmp_path / (src + ".o"))
        subprocess.run(flags + ["-fopenmp", "-x", "c++", "-c", os.path.join(CSRC, src), "-o", obj], check=True)
        objs.append(obj)
    exe = str(tmp_path / "sanitize_host")
    lib_dir = os.path.join(ROOT, "impg_b200")
    cuda_lib = os.path.join(os.path.dirname(CUDA_INC), "lib64")
    probe = subprocess.run(flags + ["-fopenmp", os.path.join(ROOT, "tests", "sanitize_host.cpp")] + objs +
                           ["-L" + lib_dir, "-limpgx", "-L" + cuda_lib, "-lcudart", "-lz", "-Wl,-rpath," + lib_dir, "-o", exe],
                           capture_output=True, text=True)
    if probe.returncode != 0 and ("asan" in probe.stderr.lower() or "cudart" in probe.stderr.lower()):
        pytest.skip("the sanitizer runtimes or libcudart are not available to the host linker")
    assert probe.returncode == 0, probe.stderr
    pafs = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.paf")))[:3]
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:protect_shadow_gap=0")
    r = subprocess.run([exe] + pafs, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.skipif(CXX is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")),
                    reason="needs g++ and the CUDA headers")
def test_text_parsers_survive_mutated_inputs_under_asan_and_ubsan(tmp_path):
    cuda_lib = os.path.join(os.path.dirname(CUDA_INC), "lib64")
    flags = [CXX, "-std=c++17", "-g", "-O1", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
             "-fno-omit-frame-pointer", "-fopenmp", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-I" + CUDA_INC]
    objs = []
    for src in ("api.cu", "partition.cu", "impg_file.cu"):
        obj = str(tmp_path / (src + ".o"))
        subprocess.run(flags + ["-x", "c++", "-c", os.path.join(CSRC, src), "-o", obj], check=True)
        objs.append(obj)
    exe = str(tmp_path / "sanitize_parsers")
    lib_dir = os.path.join(ROOT, "impg_b200")
    probe = subprocess.run(flags + [os.path.join(ROOT, "tests", "sanitize_parsers.cpp")] + objs +
                           ["-L" + lib_dir, "-limpgx", "-L" + cuda_lib, "-lcudart", "-lz", "-Wl,-rpath," + lib_dir, "-o", exe],
                           capture_output=True, text=True)
    if probe.returncode != 0 and ("asan" in probe.stderr.lower() or "cudart" in probe.stderr.lower()):
        pytest.skip("the sanitizer runtimes or libcudart are not available to the host linker")
    assert probe.returncode == 0, probe.stderr
    pafs = [os.path.join(ROOT, "tests", "golden", n) for n in ("short_floor.paf", "easy_shared_flank.paf")]
    outs = []
    for min_bytes in (str(1 << 40), "0"):  # the plain loop, then the multi-threaded PAF parse on the same inputs
        env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:protect_shadow_gap=0", IMPGX_PAF_PARALLEL_MIN_BYTES=min_bytes,
                   OMP_NUM_THREADS="4")
        r = subprocess.run([exe] + pafs, capture_output=True, text=True, env=env, timeout=900)
        assert r.returncode == 0 and r.stdout.strip().startswith("ok"), r.stdout[-2000:] + r.stderr[-4000:]
        outs.append(r.stdout.strip())
    assert outs[0] == outs[1]  # both paths accept exactly the same mutated files
