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


def test_host_input_side_matches_the_python_restatement(host_binary, golden):
    """two independent restatements of the k8s quantity grammar (C++ with 128-bit integers, Python with Fractions) agree on
    the golden annotation cases and on a few thousand generated quantity strings"""
    import random
    from oracle import pyref
    exe = os.path.join(HOST, "host_cpu_test")
    for case in golden["annotation_cases"]:
        out = subprocess.check_output([exe, "annotations"] + ["%s=%s" % kv for kv in case["annotations"].items()], text=True).strip()
        if case["error"] is not None:
            assert out == case["error"], case["id"]
        else:
            e = case["expect"]
            assert out == "ok %d %d %d %d %d %d %d %d %d" % (*e["drv"], *e["exe"], e["min"], e["max"], 1 if e["exact"] else 0), case["id"]
    rng = random.Random(20260922)
    suffixes = ["", "m", "u", "n", "k", "M", "G", "T", "P", "E", "Ki", "Mi", "Gi", "Ti", "Pi", "Ei", "e3", "E2", "e-2", "e+1", "e0",
                "Zi", "mm", "Kii", "e", "i", "K", "E-", " Gi", "x"]
    strings = []
    for _ in range(3000):
        sign = rng.choice(["", "", "", "+", "-"])
        num = rng.choice(["", "0", "00", str(rng.randrange(0, 10)), str(rng.randrange(0, 100000)), str(rng.randrange(0, 10 ** rng.randrange(1, 22)))])
        frac = rng.choice(["", "", ".", ".5", ".25", ".000", "." + str(rng.randrange(0, 10 ** rng.randrange(1, 8)))])
        strings.append(sign + num + frac + rng.choice(suffixes))
    strings += ["", "0", "1", ".", "-", "+", "1.5Gi", "2432Mi", "0.0005", "9223372036854775807", "2305843009213693951", "2305843009213693952"]
    for scale, up in ((0, 0), (3, 0), (0, 1)):
        for lo in range(0, len(strings), 200):
            chunk = strings[lo:lo + 200]
            lines = subprocess.check_output([exe, "quantity", str(scale), str(up)] + chunk, text=True).split("\n")
            for s_, line in zip(chunk, lines):
                st, v = pyref.quantity_scaled(s_, scale, round_up_fraction=bool(up))
                got_st, got_v = line.split()
                assert got_st == st, (s_, scale, up, line, st, v)
                if st == "ok":
                    assert int(got_v) == v, (s_, scale, up, line, v)
