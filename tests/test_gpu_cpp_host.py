"""Runs the C++ host-API test program (include/tnc.hpp over the C ABI) on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_binary_builds(built_lib):
    assert os.path.exists(os.path.join(ROOT, "build", "test_host_api"))


def test_cpp_hdf5_io(built_lib, tmp_path):
    """the HDF5 part of the C++ mirror (tnc::io::hdf5) needs no GPU"""
    r = subprocess.run([os.path.join(ROOT, "build", "test_host_api"), "--io", str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0 and "HOST_IO_OK" in r.stdout, r.stdout


@pytest.mark.gpu
def test_cpp_host_api(built_lib, tmp_path):
    r = subprocess.run([os.path.join(ROOT, "build", "test_host_api"), "--all", str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0 and "HOST_API_OK" in r.stdout, r.stdout
