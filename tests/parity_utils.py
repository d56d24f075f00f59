"""Shared helpers for the parity tests."""
import torch


def err_stats(got: torch.Tensor, ref: torch.Tensor):
    got, ref = got.double(), ref.double()
    err = (got - ref).abs()
    scale = ref.abs().max().clamp_min(1e-12)
    return dict(max_abs=float(err.max()), mean_abs=float(err.mean()), rel_to_max=float(err.max() / scale),
                rms_rel=float((err.pow(2).mean().sqrt()) / ref.pow(2).mean().sqrt().clamp_min(1e-12)))


def assert_fp16_close(got: torch.Tensor, ref32: torch.Tensor, what: str, rtol: float = 1e-3, atol_frac: float = 1e-3):
    """fp16 result vs an fp32 reference of the same op on the same fp16 inputs.

    Tolerance = north_star's rtol 1e-3 (one fp16 rounding is <= 4.9e-4 relative) plus an absolute term of
    atol_frac * max|ref| for elements that are small only through cancellation (their error scales with the operands,
    not with the result).  north_star's literal atol 1e-4 is below fp16 resolution for |x| > 0.2 and is applied to the
    DDIM step, which is checked bit-exactly instead.
    """
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    ref32 = ref32.float()
    atol = atol_frac * float(ref32.abs().max())
    bad = (got.float() - ref32).abs() > (atol + rtol * ref32.abs())
    frac = float(bad.float().mean())
    assert frac == 0.0, f"{what}: {frac:.2e} of elements outside rtol={rtol} atol={atol:.2e}; stats {err_stats(got, ref32)}"
