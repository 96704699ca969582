"""The C++ host-side mirror of the reference's plug-in interface
(k8s-spark-scheduler_b200/host/gangpack_host.hpp; tests in tests/host_test.cpp read like the
reference's own Go tests).  CPU: it compiles and links against libgangpack.so and refuses to run
without a device.  GPU: the suite runs through the C ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "k8s-spark-scheduler_b200", "host")
BIN = os.path.join(HOST, "host_test")


@pytest.fixture(scope="module")
def host_binary():
    import k8s_spark_scheduler_b200 as g
    g.native.build()
    subprocess.check_call(["make", "-C", HOST, "-s"])
    return BIN


def test_host_layer_builds_and_fails_loudly_without_gpu(host_binary):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    p = subprocess.run([host_binary], capture_output=True, text=True)
    assert p.returncode != 0
    assert "no CPU path" in (p.stderr + p.stdout)


@pytest.mark.gpu
def test_host_layer_suite(host_binary):
    p = subprocess.run([host_binary], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "all passed" in p.stdout


def test_host_input_side_on_cpu(host_binary):
    """driver annotations -> application tuple (sparkResources + the k8s quantity grammar), the reference's own
    sparkpods_test.go cases; no device involved"""
    p = subprocess.run([os.path.join(HOST, "host_cpu_test")], capture_output=True, text=True, timeout=60)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "all passed" in p.stdout
