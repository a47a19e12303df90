"""-m gpu: every C-ABI kernel configuration against a plain PyTorch fp32 computation of the same
operator (tolerances are set per case in kernel_cases.py: fp32 outputs 2e-4 absolute on O(1) values,
bf16 outputs one bf16 rounding, bf16x3 split mode 1e-4)."""
import pytest

import kernel_cases as kc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(kc.CASES.keys()))
def test_kernel_case(name):
    from sscg_b200 import kernels as K
    err, scale, tol = kc.CASES[name]()
    assert K.device_error() == 0
    assert err <= tol, f"{name}: max abs err {err:.3e} > tol {tol:.3e} (ref scale {scale:.3e})"
