"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol include/nuts_b200.h declares,
fails loudly without a GPU (no CPU fallback), and the ctypes mirror matches the C structs."""
import ctypes as C
import os
import re
import subprocess

import pytest

from nuts_rs_b200 import _abi, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "nuts_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nuts_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    if not os.path.exists(lib.LIB_PATH):
        pytest.skip("libnuts_b200.so not built (run __graft_entry__.build())")
    L = C.CDLL(lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 50
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(lib.EXPORTED_SYMBOLS) == declared


def test_struct_layout_matches_c():
    """Compile a tiny C program printing sizeof/offsetof of the ABI structs and compare with ctypes."""
    prog = r"""
    #include <stdio.h>
    #include <stddef.h>
    #include "nuts_b200.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(nuts_settings_t), offsetof(nuts_settings_t, adapt_options),
             offsetof(nuts_settings_t, check_turning), offsetof(nuts_settings_t, seed), sizeof(nuts_logp_desc_t),
             sizeof(nuts_stats_t), offsetof(nuts_euclidean_adapt_options_t, early_window), sizeof(nuts_step_size_settings_t));
      return 0;
    }"""
    import tempfile

    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(td, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        got = list(map(int, subprocess.check_output([exe]).split()))
    want = [C.sizeof(_abi.NutsSettings), _abi.NutsSettings.adapt_options.offset, _abi.NutsSettings.check_turning.offset,
            _abi.NutsSettings.seed.offset, C.sizeof(_abi.LogpDesc), C.sizeof(_abi.Stats), _abi.EuclideanAdaptOptions.early_window.offset,
            C.sizeof(_abi.StepSizeSettings)]
    assert got == want


def test_settings_default_matches_reference_defaults():
    if not os.path.exists(lib.LIB_PATH):
        pytest.skip("libnuts_b200.so not built")
    s = lib.DiagNutsSettings()
    p = _abi.default_settings()
    assert bytes(s) == bytes(p)
    # reference src/sampler.rs:507-531,630-634 ; src/adapt_strategy.rs:56-69 ; src/stepsize/adapt.rs:320-329 ; dual_avg.rs:22-31
    assert (s.num_tune, s.num_draws, s.maxdepth, s.mindepth, s.num_chains, s.seed) == (400, 1000, 10, 0, 6, 0)
    assert s.max_energy_error == 1000.0 and s.check_turning == 1 and s.extra_doublings == 0
    a = s.adapt_options
    assert (a.early_window, a.step_size_window, a.mass_matrix_switch_freq, a.early_mass_matrix_switch_freq) == (0.3, 0.15, 80, 10)
    assert a.mass_matrix_update_freq == 1 and a.mass_matrix_window_growth == 1.5
    assert a.mass_matrix_options.use_grad_based_estimate == 1
    ss = a.step_size_settings
    assert (ss.target_accept, ss.initial_step, ss.has_jitter, ss.jitter) == (0.8, 0.1, 1, 0.1)
    da = ss.adapt_options.dual_average
    assert (da.k, da.t0, da.gamma) == (0.75, 10.0, 0.05) and abs(da.max_step_size - 3.141592653589793) < 1e-15


def test_no_cpu_fallback():
    """Without a GPU every compute entry point must fail loudly instead of silently computing on the host."""
    if not os.path.exists(lib.LIB_PATH):
        pytest.skip("libnuts_b200.so not built")
    if lib.device_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lib.NutsError, match="no CUDA device|CPU fallback|sm_"):
        lib.CudaMath(4, 10, _abi.NUTS_LOGP_GAUSS_ISO, mu=3.0)
