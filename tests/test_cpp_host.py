"""The C++ host side (include/nuts_b200.hpp) over the C ABI: the reference is compiled code, so its public surface for this path
(DiagNutsSettings / Model+Math / Chain, reference benches/sample.rs:79-99) is mirrored in C++; this test builds the example the way
a user would (g++ -Iinclude ... -lnuts_b200) and checks
  * without a GPU: it links against every entry point it uses and fails loudly with NUTS_ERR_NO_DEVICE (exit code 77);
  * on the GPU (-m gpu): it samples N(3, I), passes its own moment checks, and its draws are the same numbers the Python mirror
    gets from the same library for the same settings and seed (checksum over every coordinate)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path_factory, name):
    out = tmp_path_factory.mktemp("cpp") / name
    libdir = os.path.join(ROOT, "nuts_rs_b200")
    cmd = ["g++", "-std=c++17", "-pthread", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", name + ".cpp"), "-L", libdir, "-lnuts_b200", "-Wl,-rpath," + libdir, "-o", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return str(out)


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    return _build(tmp_path_factory, "sample_normal")


@pytest.fixture(scope="module")
def exe_control(tmp_path_factory):
    return _build(tmp_path_factory, "sampler_control")


def test_cpp_example_builds_and_has_no_cpu_fallback(exe):
    from nuts_rs_b200 import lib

    if lib.device_available():
        pytest.skip("a device is visible: covered by the gpu test")
    r = subprocess.run([exe, "4", "10", "20", "20"], capture_output=True, text=True)
    assert r.returncode == 77, (r.returncode, r.stderr)
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_example_matches_python_mirror(exe):
    from nuts_rs_b200 import _abi, lib

    N, d, tune, n = 6, 10, 200, 300
    r = subprocess.run([exe, str(N), str(d), str(tune), str(n)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    fields = r.stdout.split()
    got = {fields[i]: fields[i + 1] for i in range(0, len(fields), 2)}
    assert int(got["direct"]) == 1  # Draws are page-locked: written by the kernel itself
    m = lib.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_ISO, mu=3.0)
    s = lib.Sampler(m, lib.DiagNutsSettings(num_tune=tune, maxdepth=3), seed=42)
    assert (s.set_position(np.full((N, d), 3.5)) == 0).all()
    s.draw(tune)
    draws, stats = s.draw(n)
    s.close()
    m.close()
    t, c, i = np.meshgrid(np.arange(n), np.arange(N), np.arange(d), indexing="ij")
    w = 1 + (i + 3 * c + 7 * t) % 11
    # same accumulation order as the C++ loop (draw, chain, coordinate)
    checksum = 0.0
    for v in (draws * w).ravel():
        checksum += v
    assert float(got["checksum"]) == checksum
    assert int(got["leapfrogs"]) == int(stats["n_steps"].sum())


def test_cpp_sampler_control_builds_and_has_no_cpu_fallback(exe_control):
    from nuts_rs_b200 import lib

    if lib.device_available():
        pytest.skip("a device is visible: covered by the gpu test")
    r = subprocess.run([exe_control], capture_output=True, text=True)
    assert r.returncode == 77, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_cpp_sampler_pause_resume_progress_checkpoint(exe_control):
    """nuts_b200::Sampler (the reference's Sampler::pause / resume / progress / wait, src/sampler.rs:1253-1552, at batch
    granularity), the init-retry loop (src/sampler.rs:1133-1143) and checkpoint / restore through the C ABI."""
    r = subprocess.run([exe_control], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    fields = r.stdout.split()
    got = {fields[i]: fields[i + 1] for i in range(0, len(fields), 2)}
    assert int(got["resumed_identical"]) == 1 and int(got["leapfrogs"]) > 0
