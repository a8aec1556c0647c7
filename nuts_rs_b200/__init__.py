"""nuts_rs_b200 — B200-native many-chain NUTS hot path behind a C ABI (libnuts_b200.so).

Importing the package never needs a GPU; every compute entry point fails loudly (no CPU fallback) when the
CUDA library or an sm_100 device is missing.
"""
from . import _abi  # noqa: F401
from ._abi import (  # noqa: F401
    NUTS_LOGP_FUNNEL,
    NUTS_LOGP_GAUSS_DIAG,
    NUTS_LOGP_GAUSS_ISO,
    NUTS_LOGP_GAUSS_RANK1,
    NUTS_STEPSIZE_DUAL_AVERAGE,
    NUTS_STEPSIZE_FIXED,
    default_settings,
)
