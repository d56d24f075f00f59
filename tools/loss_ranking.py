"""Rank the kernel shapes of the measured step profile (profiles/r02_step_profile.txt, in-situ CUDA-event times per op and
shape) by the time they lose against their own roofline: max(FLOPs / 1430 TF/s sustained, algorithmic bytes / 6.57 TB/s).
Only the shapes the profile lists (top 30 per phase) are covered.    python tools/loss_ranking.py"""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TF, BW = 1430.5e12, 6570e9


def sol_us(kind, kv):
    g = lambda k: int(kv[k])
    if kind.startswith(("linear", "conv3x3", "tconv3")):
        M, N, K = g("M"), g("N"), g("K")
        fl = 2.0 * M * N * K
        a_bytes = M * K * 2 if kind.startswith("linear") else M * (K // 9) * 2 if kind.startswith("conv3x3") else M * (K // 3) * 2
        out = M * (N // 2 if "geglu" in kind else N) * 2
        by = a_bytes + N * K * 2 + out * (2 if "+res" in kind else 1)
        return max(fl / TF, by / BW) * 1e6
    if kind.startswith("attention"):
        batch, seq, heads, nv = g("batch"), g("seq"), g("heads"), g("nv")
        if "frames" in kind:  # batch = clips*HW sequences of `seq` frames
            fl = 4.0 * batch * heads * seq * seq * 64
            by = 4.0 * batch * seq * heads * 64 * 2
        else:
            fl = 2.0 * batch * heads * seq * seq * 64 * (1 + nv)
            by = (2.0 + 2 * nv) * batch * seq * heads * 64 * 2
        return max(fl / TF, by / BW) * 1e6
    if kind.startswith("groupnorm"):
        return 4.0 * g("n") * g("rows") * g("C") / BW * 1e6
    if kind.startswith("layernorm"):
        return 4.0 * g("rows") * g("C") / BW * 1e6
    return None


def main():
    rows, phase = [], None
    for line in open(os.path.join(ROOT, "profiles", "r02_step_profile.txt")):
        if line.startswith("==="):
            phase = "inversion" if "inversion" in line else "edit"
            continue
        m = re.match(r"\s+([\d.]+) ms n=\s*(\d+) avg\s+([\d.]+) us\s+(.*)", line)
        if not m:
            continue
        total_ms, n, avg, desc = float(m.group(1)), int(m.group(2)), float(m.group(3)), m.group(4).strip()
        kind = desc.split(" M=")[0].split(" n=")[0].split(" rows=")[0].split(" nv=")[0].strip()
        if desc.startswith("attention"):
            kind = " ".join(desc.split()[:2])
        kv = dict(re.findall(r"(\w+)=(\d+)", desc))
        s = sol_us(kind if not desc.startswith(("linear", "conv3x3", "tconv3")) else desc.split()[0], kv)
        if s is None:
            continue
        rows.append((phase, (avg - s) * n / 1e3, desc, n, avg, s))
    for phase in ("edit", "inversion"):
        sel = sorted([r for r in rows if r[0] == phase], key=lambda r: -r[1])
        print(f"== {phase} step: time lost against the per-shape roofline (listed shapes only), ms per step")
        print(f"{'lost ms':>8s} {'calls':>5s} {'measured us':>11s} {'roofline us':>11s}  {'eff':>5s}  shape")
        for _, lost, desc, n, avg, s in sel[:18]:
            print(f"{lost:8.2f} {n:5d} {avg:11.1f} {s:11.1f}  {s / avg * 100:4.0f}%  {desc}")
        print(f"   sum over the listed shapes: {sum(r[1] for r in sel):.1f} ms")


if __name__ == "__main__":
    main()
