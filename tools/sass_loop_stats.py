"""Static check of the HBM-bound kernels' inner loops (no GPU needed): instructions issued per 16-byte vector against the
issue budget the HBM roofline leaves.

B200: 148 SMs x 4 schedulers x 1 warp-instruction / clk = 128 thread-instructions / clk / SM.  At the measured 6.57 TB/s and
~1.9 GHz an SM has to move 6570e9 / 148 / 1.9e9 = 23 B / clk to keep up, i.e. one 16-byte vector per 0.69 clk per SM ->
128 * 0.69 = 88 thread-instructions per vector for a read-only pass and 176 per vector for a pass that reads one vector and
writes one (bytes double, time doubles).  A loop that needs more than that is instruction-issue bound, not memory bound,
whatever its access pattern; one that needs 70-80 % of it has no slack left to hide latency.  The count is static: a loop
that holds both sides of a warp-uniform branch (gn_apply: SiLU / no SiLU) is counted twice.  Usage: python tools/sass_loop_stats.py [object file]
"""
import re
import subprocess
import sys

HBM, SMS, GHZ = 6570e9, 148, 1.9e9


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    cur, body = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            body[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return body


def innermost_loops(instrs):
    """backward branches -> (start, end) address ranges; keep the ones that contain no other loop"""
    loops = []
    for addr, text in instrs:
        m = re.search(r"\bBRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", text)
        if m and int(m.group(1), 16) <= addr:
            loops.append((int(m.group(1), 16), addr))
    return [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]


def report(obj, patterns):
    for name, instrs in functions(obj).items():
        if not any(p in name for p in patterns):
            continue
        for lo, hi in innermost_loops(instrs):
            body = [t for a, t in instrs if lo <= a <= hi]
            ld = sum(1 for t in body if re.search(r"\bLDG\.E\.128|LDGSTS", t))
            stv = sum(1 for t in body if re.search(r"\bSTG\.E\.128", t))
            if ld < 2:
                continue
            mufu = sum(1 for t in body if "MUFU" in t)
            per_vec = len(body) / ld
            bytes_per_vec = 16 * (1 + (stv / ld))
            budget = 128 * bytes_per_vec / (HBM / SMS / GHZ)
            clk_per_vec = bytes_per_vec / (HBM / SMS / GHZ)          # time the roofline leaves per thread-vector, per SM
            mufu_budget = 16 * clk_per_vec                            # 16 MUFU lanes / clk / SM
            print(f"{name[:90]:90s} loop {lo:#06x}-{hi:#06x}: {len(body):4d} instr (static; both sides of uniform branches), "
                  f"{ld} x 16 B loads, {stv} x 16 B stores, {mufu} MUFU -> {per_vec:5.1f} instr / vector of a {budget:4.0f} budget "
                  f"({per_vec / budget * 100:3.0f} %), {mufu / ld:4.1f} MUFU / vector of {mufu_budget:4.1f} ({mufu / ld / mufu_budget * 100:3.0f} %)")


if __name__ == "__main__":
    obj = sys.argv[1] if len(sys.argv) > 1 else "/tmp/elementwise.o"
    report(obj, ("gn_", "layernorm", "ddim"))
