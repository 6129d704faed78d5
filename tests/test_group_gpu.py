"""-m gpu: the native multi-GPU entry point (jpezyb200_group_create / jpezyb200_group_encode, capi_group.inc) driven by a C++
host program with no Python in the loop (tests/host/group_encode_check.cpp): the stitched segment must be byte-identical to
the single-GPU segment.  Ranks are emulated on device 0 (the same device named several times) and, where the box has
more than one GPU, spread over the real devices (peer stores over NVLink)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "group_encode_check")


@pytest.fixture(scope="module")
def exe():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(cuda, "include"), "-o", EXE,
                           os.path.join(ROOT, "tests", "host", "group_encode_check.cpp"), "-L" + os.path.join(ROOT, "jpezy_b200"),
                           "-ljpezy_b200", "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath," + os.path.join(ROOT, "jpezy_b200")])
    return EXE


@pytest.mark.parametrize("W,H,family,gray,ranks", [(64, 48, 0, 0, 2), (208, 128, 1, 0, 3), (333, 77, 0, 1, 2), (640, 360, 1, 0, 8),
                                                   (1920, 1080, 0, 0, 4), (4096, 1024, 1, 0, 8), (48, 16, 2, 0, 1)])
def test_group_encode_from_a_cpp_host_emulated_ranks(exe, W, H, family, gray, ranks):
    p = subprocess.run([exe, str(W), str(H), str(family), str(gray)] + ["0"] * ranks, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "identical to the single-GPU segment" in p.stdout


def test_group_encode_across_real_devices(exe):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU on this box")
    p = subprocess.run([exe, "8192", "4096", "0", "0"] + [str(k) for k in range(n)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
