"""Numerical emulation (torch, CPU) of the per-row algorithm of csrc/attention2q_tcgen05.cu: 128-key tiles, running max
raised only when a tile exceeds it by > 2^8 (log2 domain) with O / l rescaled, P rounded to fp16 before the PV product,
fp32 accumulation, 0 / 25 / 50 % of the exponentials through the FMA-pipe polynomial, key-tail masking.  Checks the
ALGORITHM (not the hardware protocol) against exact softmax attention at the tolerance of the GPU parity tests."""
import numpy as np
import torch

from tools import exp2_poly_fit

TK, THRESH = 128, 8.0


def _ex2(a: torch.Tensor, poly_mask: torch.Tensor) -> torch.Tensor:
    exact = torch.exp2(a)
    if not poly_mask.any():
        return exact
    p = torch.from_numpy(exp2_poly_fit.ex2_poly(a.numpy().astype(np.float32)).astype(np.float32))
    return torch.where(poly_mask, p, exact)


def attention_2q_rows(q, k, v, scale=0.125, poly=0):
    """q [T,64], k/v [L,64] fp16 -> [T,64] fp16, one head"""
    T, L = q.shape[0], k.shape[0]
    sc = np.float32(scale * 1.4426950408889634)
    qf, kf, vf = q.float(), k.float(), v.float()
    m = torch.zeros(T)
    l = torch.zeros(T)
    o = torch.zeros(T, 64)
    # element e of a 32-column chunk (pairs (2e, 2e+1)) goes to the polynomial when (e & 3) == 3 (25 %) / (e & 1) (50 %)
    col = torch.arange(TK)
    e = (col % 32) // 2
    pm = ((e & 3) == 3) if poly == 1 else ((e & 1) == 1) if poly == 2 else torch.zeros(TK, dtype=torch.bool)
    for j in range((L + TK - 1) // TK):
        kt, vt = kf[j * TK:(j + 1) * TK], vf[j * TK:(j + 1) * TK]
        n = kt.shape[0]
        s = torch.full((T, TK), float("-inf"))
        s[:, :n] = qf @ kt.T                     # fp32 accumulate of fp16 products (tensor core)
        rmax = s.max(dim=1).values * sc
        if j == 0:
            m_new = torch.where(torch.isinf(rmax), torch.zeros_like(rmax), rmax)
        else:
            need = rmax > m + THRESH
            m_new = torch.where(need, rmax, m)
            f = torch.where(need, torch.exp2(m - m_new), torch.ones_like(m))
            o, l = o * f[:, None], l * f
        m = m_new
        a = s * sc - m[:, None]
        p = _ex2(a.float(), pm[None, :].expand(T, TK))
        l = l + p.sum(dim=1)
        p16 = p.half().float()
        o = o + p16[:, :n] @ vt
    return (o / l[:, None]).half()


def check(T=192, L=880, mag=2.0, poly=2, seed=0, rising=False):
    g = torch.Generator().manual_seed(seed)
    q = (torch.randn(T, 64, generator=g) * mag).half()
    k = torch.randn(L, 64, generator=g) * mag
    if rising:
        k = k * torch.linspace(0.2, 6.0, L)[:, None]
    k = k.half()
    v = torch.randn(L, 64, generator=g).half()
    got = attention_2q_rows(q, k, v, poly=poly).float()
    ref = torch.softmax(q.double() @ k.double().T * 0.125, dim=-1) @ v.double()
    err = (got.double() - ref).abs()
    tol = 2e-3 * ref.abs().max() + 1e-3 * ref.abs()
    return float((err / tol).max())


if __name__ == "__main__":
    for poly in (0, 1, 2):
        for kw in (dict(), dict(L=145, T=256), dict(mag=6.0), dict(rising=True, L=1024)):
            print(poly, kw, f"worst error / tolerance = {check(poly=poly, **kw):.3f}")
