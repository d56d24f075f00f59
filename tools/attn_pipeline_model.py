"""Cycle-level what-if model of the attention pipelines (no GPU): v9 (shipped), v10 (AV2V_ATTN_V10) and the two-query-tile
kernel (AV2V_ATTN_2Q), built on the event loop of tools/protocol_sim.py with FIXED latencies instead of random ones.

Resources: one in-order tensor pipe per SM; one MUFU port per SM sub-partition shared by the softmax warps that live on it
(group A's and group B's warp of the same lane quarter).  Numbers from profiles/r01_attention_phase_timers.txt: inside the
softmax instruction mix one warp alone needs 12.3 cycles per MUFU instruction (it cannot saturate the port), two warps together
get one instruction per 9.0 cycles -> the ex2 pass of 128 scores per row is modelled as 16 slices of 8 instructions, each
needing 98 cycles of the warp's own time and 72 cycles of the shared port (FIFO).  tcgen05.ld + wait + row max 330, max
hand-over 170, P store + fences 90, QK^T 256, PV 256 * NV tensor cycles, per-item prologue / epilogue 144 per tile, and a
fitted 80 cycles of signalling latency per mbarrier hand-off.
The model is checked against v9 (measured ~1900-2200 cycles per 128 x 128 tile for NV = 1 and NV = 3) and then asked what the
other pipelines would do.  It ignores second-order effects (issue contention with the poll loops, smem bandwidth, clocks).

    python tools/attn_pipeline_model.py
"""
from __future__ import annotations

import random
import sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from protocol_sim import Barrier, Chan, Sim  # noqa: E402

LD, MAX, HANDOVER, STORE, S_MMA, PV_UNIT = 150, 180, 170, 90, 256, 256
N_SLICES, SLICE_OWN, SLICE_PORT, ITEM_OVERHEAD = 16, 98, 72, 144
HOP = 80  # mbarrier / tcgen05.commit signalling latency per hand-off (fitted: brings the v9 model to the measured ~2000 cycles)


class Mufu:
    def __init__(self):
        self.free_at = 0

    def run(self, sim, mufu_frac, extra_own):
        """generator: the ex2 pass of one key tile; `mufu_frac` of the exponentials on MUFU, the rest (FMA-pipe polynomial)
        only cost the warp's own issue time (`extra_own` cycles per slice)"""
        for _ in range(N_SLICES):
            t0 = sim.t
            start = max(sim.t, self.free_at)
            self.free_at = start + SLICE_PORT * mufu_frac
            finish = max(self.free_at, t0 + SLICE_OWN * mufu_frac + extra_own)
            yield ("delay", finish - sim.t)


def model(kind, nv, n_kv=32, items=3, mufu_frac=1.0, extra_own=0.0):
    """-> cycles per 128 x 128 (query tile x key tile) in steady state"""
    sim = Sim(random.Random(0))
    one = lambda n: Barrier(n, 1)
    s_full, p_ready, pv_done = [one("s_full0"), one("s_full1")], [Barrier("p_ready0", 1), Barrier("p_ready1", 1)], [one("pv0"), one("pv1")]
    s_free = [one("s_free0"), one("s_free1")]
    mufu = Mufu()
    chan = [Chan("c0"), Chan("c1")]
    done = {"tiles": 0, "t_first": None, "t_last": 0}
    pv_cycles = PV_UNIT * nv
    two_q = kind == "2q"

    def mma():
        g = 0
        if two_q:
            for it in range(items):
                for j in range(n_kv):
                    if j == 0:
                        for x in range(2):
                            if g > 0:
                                yield ("wait", s_free[x], (g - 1) & 1)
                                yield ("delay", HOP)
                            sim.mma(S_MMA, lambda: None)
                            sim.commit(s_full[x])
                    if j + 1 < n_kv:
                        for x in range(2):
                            yield ("wait", s_free[x], g & 1)
                            yield ("delay", HOP)
                            sim.mma(S_MMA, lambda: None)
                            sim.commit(s_full[x])
                    for x in range(2):
                        yield ("wait", p_ready[x], g & 1)
                        yield ("delay", HOP)
                        sim.mma(pv_cycles, lambda: None)
                        sim.commit(pv_done[x])
                    g += 1
            return
        for it in range(items):
            def issue_s(gg):
                if kind == "v10" and gg >= 2:
                    yield ("wait", s_free[gg & 1], ((gg - 2) >> 1) & 1)
                    yield ("delay", HOP)
                sim.mma(S_MMA, lambda: None)
                sim.commit(s_full[gg & 1])

            def issue_pv(gg):
                yield ("wait", p_ready[gg & 1], (gg >> 1) & 1)
                yield ("delay", HOP)
                sim.mma(pv_cycles, lambda: None)
                sim.commit(pv_done[gg & 1])
            yield from issue_s(g)
            yield from issue_s(g + 1)
            for j in range(n_kv):
                if kind == "v9":
                    yield from issue_pv(g)
                    if j + 2 < n_kv:
                        yield from issue_s(g + 2)
                else:
                    if j + 2 < n_kv:
                        yield from issue_s(g + 2)
                    if j >= 1:
                        yield from issue_pv(g - 1)
                g += 1
            if kind == "v10":
                yield from issue_pv(g - 1)

    def softmax(grp):
        g = 0
        n = 0
        for it in range(items):
            for j in range(n_kv):
                if not two_q and (g & 1) != grp:
                    g += 1
                    continue
                cnt = n if two_q else g
                yield ("wait", s_full[grp], (cnt & 1) if two_q else ((g >> 1) & 1))
                yield ("delay", HOP)
                yield ("delay", LD)
                if two_q or kind == "v10":
                    s_free[grp].arrive()
                yield ("delay", MAX)
                if not two_q:
                    if j > 0:
                        yield ("wait", chan[grp], 0)
                        yield ("delay", HOP)
                        chan[grp].get()
                    chan[grp ^ 1].put(1)
                    yield ("delay", HANDOVER)
                if two_q and n > 0:
                    yield ("wait", pv_done[grp], (n - 1) & 1)
                    yield ("delay", HOP)
                yield from mufu.run(sim, mufu_frac, extra_own)
                if kind == "v10" and g > 0:
                    yield ("wait", pv_done[grp ^ 1], ((g - 1) >> 1) & 1)
                    yield ("delay", HOP)
                yield ("delay", STORE)
                p_ready[grp].arrive()
                done["tiles"] += 1
                if done["t_first"] is None and it == 1:
                    done["t_first"], done["n_first"] = sim.t, done["tiles"]
                done["t_last"], done["n_last"] = sim.t, done["tiles"]
                g += 1
                n += 1
            if not two_q:  # epilogue hand-over of the row sums (both groups)
                gl = g - 1
                if (gl & 1) != grp:
                    yield ("wait", chan[grp], 0)
                    yield ("delay", HOP)
                    chan[grp].get()
                chan[grp ^ 1].put(1)
                yield ("wait", chan[grp], 0)
                yield ("delay", HOP)
                chan[grp].get()
            yield ("delay", ITEM_OVERHEAD * n_kv // (1 if two_q else 2))  # item prologue / epilogue (O -> global), per tile it processed

    sim.spawn("mma", mma())
    for grp in range(2):
        sim.spawn(f"softmax{grp}", softmax(grp))
    sim.run()
    return (done["t_last"] - done["t_first"]) / max(1, done["n_last"] - done["n_first"])


if __name__ == "__main__":
    print("cycles per 128 x 128 tile (query tile x key tile), 32 key tiles per item; measured v9: ~1900-2100 for NV = 1 and NV = 3")
    for nv in (1, 3):
        row = [f"NV={nv}:"]
        row.append(f"v9 {model('v9', nv):6.0f}")
        row.append(f"v10 {model('v10', nv):6.0f}")
        # 3 of 8 key pairs through the packed polynomial: 5/8 of the MUFU work; the polynomial costs ~12 issue slots per pair ->
        # 48 pairs x 12 / 16 slices = 36 cycles of the warp's own time per slice
        row.append(f"v10 + 3/8 FMA-pipe exp2 (packed) {model('v10', nv, mufu_frac=0.625, extra_own=36):6.0f}")
        if nv == 1:
            row.append(f"2q {model('2q', nv):6.0f}")
            row.append(f"2q + 3/8 FMA-pipe exp2 (packed) {model('2q', nv, mufu_frac=0.625, extra_own=36):6.0f}")
        print("  ".join(row))
