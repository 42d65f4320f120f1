"""Builds and runs the C++ host-mirror test (tests/cpp/test_barcode_matching.cpp): the reference's own assign()
test cases, written against include/fqtk_b200.hpp, linked straight to the C-ABI library."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "fqtk_b200")


def _build(tmp_path):
    exe = str(tmp_path / "test_barcode_matching")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_barcode_matching.cpp"), "-o", exe,
           "-L", LIBDIR, "-lfqtk_b200", f"-Wl,-rpath,{LIBDIR}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_cpp_mirror_compiles_against_the_c_abi(tmp_path):
    assert os.path.exists(_build(tmp_path))


@pytest.mark.gpu
def test_cpp_mirror_reference_cases(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all reference assign cases pass" in r.stdout


def test_compressed_key_primitives_on_cpu(tmp_path):
    """acgt_key, acgt_only, nibble_distance, g4_hashes (csrc/common.cuh, shared by host and device): valid <=> every nibble one-hot, and injective
    on valid reads — checked on the CPU (tests/cpp/test_keys.cpp), no GPU needed."""
    exe = str(tmp_path / "test_keys")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-I", "/usr/local/cuda/include",
           os.path.join(ROOT, "tests", "cpp", "test_keys.cpp"), "-o", exe]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "injective on valid reads" in r.stdout


def test_cpp_host_mirror_formatting_on_cpu(tmp_path):
    """write_header / demux_metrics of include/fqtk_b200.hpp against the reference's own tests — no GPU needed (the
    library is linked only because the header also declares the matcher)."""
    exe = str(tmp_path / "test_host_mirror")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp"), "-o", exe,
           "-L", LIBDIR, "-lfqtk_b200", f"-Wl,-rpath,{LIBDIR}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "match the reference's tests" in r.stdout


def _build_bgzf(tmp_path):
    exe = str(tmp_path / "test_bgzf")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_bgzf.cpp"), "-o", exe,
           "-L", LIBDIR, "-lfqtk_b200", f"-Wl,-rpath,{LIBDIR}", "-lz"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_cpp_bgzf_writer_compiles_against_the_c_abi(tmp_path):
    assert os.path.exists(_build_bgzf(tmp_path))


@pytest.mark.gpu
def test_cpp_bgzf_writer_members_inflate_with_zlib(tmp_path):
    """BgzfPool / BgzfWriter (include/fqtk_b200.hpp, the mirror of the reference's pooled BGZF writers) checked by zlib's
    inflater from C++: no Python in the loop."""
    r = subprocess.run([_build_bgzf(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "every member inflates" in r.stdout
    assert "every run inflates to its records" in r.stdout  # fqtk_b200_demux_chunks, the one-call batch form
