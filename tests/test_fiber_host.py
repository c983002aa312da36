"""The fiber switch of the batched multi-chain driver (stan_b200/cpp/b200/fiber.hpp), built for the host and run
without a GPU: 1024 chains' worth of fibers in one thread, yields from inside a recursion, exceptions thrown and
caught in one fiber while others are suspended inside try blocks."""
import ctypes as C
import math
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("fiber_host") / "libfiber_host.so"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I", os.path.join(ROOT, "stan_b200", "cpp"),
                    "-o", str(out), os.path.join(ROOT, "tests", "host", "fiber_host.cpp")], check=True)
    return C.CDLL(str(out))


def reference(n_fibers, depth):
    """the same computation without fibers"""
    tot_leaves = tot_value = 0
    fp = 0.0
    for i in range(n_fibers):
        leaves = 0

        def build(d):
            nonlocal leaves
            if d == 0:
                leaves += 1
                if (leaves + i) % 97 == 0:
                    raise ValueError
                return 1
            n = 0
            try:
                n += build(d - 1)
                n += build(d - 1)
            except ValueError:
                n += 1000000
            return n
        for rep in range(3):
            tot_value += build(depth)
            fp += math.sqrt(i + rep + 1)
        tot_leaves += leaves
    return tot_leaves, tot_value, fp


@pytest.mark.parametrize("n_fibers,depth", [(1, 0), (7, 3), (1024, 6), (64, 10)])
def test_fibers_interleave_like_sequential_runs(lib, n_fibers, depth):
    leaves, value, fp = C.c_long(), C.c_long(), C.c_double()
    sweeps = lib.fiber_selftest(n_fibers, depth, C.byref(leaves), C.byref(value), C.byref(fp))
    r_leaves, r_value, r_fp = reference(n_fibers, depth)
    assert (leaves.value, value.value) == (r_leaves, r_value)
    # each fiber accumulates its own sum and the totals are added in the same order: bitwise equal
    fp_ref = 0.0
    for i in range(n_fibers):
        acc = 0.0
        for rep in range(3):
            acc += math.sqrt(i + rep + 1)
        fp_ref += acc
    assert fp.value == fp_ref
    assert sweeps >= 3       # lock-step: at least one sweep per yield of the longest fiber
