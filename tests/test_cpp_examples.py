"""The reference's only known-answer test (examples/src/modes_of_use/common.cuh:29-105), written against
this library's C++ wrapper API in the reference's three modes of use: examples/modes_of_use.cpp.
CPU: it compiles with plain g++ and links against the shared library.  GPU: it reproduces the answer."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BINARY = os.path.join(ROOT, "examples", "bin", "modes_of_use")


def test_examples_build_with_a_host_compiler_only(library):
    subprocess.run(["make", "-C", ROOT, "examples"], check=True, capture_output=True)
    assert os.path.exists(BINARY)
    needed = subprocess.run(["ldd", BINARY], capture_output=True, text=True, check=True).stdout
    assert "libgiddy_b200.so" in needed and "not found" not in needed.split("libgiddy_b200.so")[1].splitlines()[0]


@pytest.mark.gpu
def test_three_modes_of_use_reproduce_the_known_answer(library):
    if not os.path.exists(BINARY):
        subprocess.run(["make", "-C", ROOT, "examples"], check=True, capture_output=True)
    result = subprocess.run([BINARY], capture_output=True, text=True)
    assert result.returncode == 0, result.stdout + result.stderr
    assert result.stdout.count("SUCCESS") == 3, result.stdout


@pytest.mark.gpu
def test_sharded_scan_over_all_visible_gpus(library):
    """examples/sharded_scan.cpp: one column cut into segment-range shards by the library's planner, one per GPU (worker
    threads of a device group), the fused decoder and the two-kernel chain on every shard, a dictionary broadcast."""
    binary = os.path.join(ROOT, "examples", "bin", "sharded_scan")
    if not os.path.exists(binary):
        subprocess.run(["make", "-C", ROOT, "examples"], check=True, capture_output=True)
    result = subprocess.run([binary, "22"], capture_output=True, text=True, timeout=600)
    assert result.returncode == 0, result.stdout + result.stderr
    assert result.stdout.count("SUCCESS") == 3, result.stdout


@pytest.mark.gpu
def test_compress_and_decompress_through_the_factory(library):
    """examples/compress_roundtrip.cpp: the compressors' factory keys used from C++ in the reference's mode (iii) idiom -- RLE
    there and back, and Delta + DZB-variable compressed by two kernels and decompressed by the one fused decoder."""
    binary = os.path.join(ROOT, "examples", "bin", "compress_roundtrip")
    if not os.path.exists(binary):
        subprocess.run(["make", "-C", ROOT, "examples"], check=True, capture_output=True)
    result = subprocess.run([binary, "22"], capture_output=True, text=True, timeout=600)
    assert result.returncode == 0, result.stdout + result.stderr
    assert result.stdout.count("SUCCESS") == 2, result.stdout
