"""Discrete-event models of the mbarrier protocols of csrc/attention2q_tcgen05.cu and csrc/attention_tfused_tcgen05.cu (one- and
two-slot kernels) — CPU-only verification aid.

The kernel's warp roles are restated as Python generators that perform the SAME sequence of mbarrier waits / arrives /
tcgen05.commit / TMA operations with the SAME parity expressions; an event loop runs them with randomised latencies.
Every buffer (Q / K / V smem stages, S / P / O TMEM regions) carries a data tag and a reader/writer state, and each
access asserts that it sees exactly the data the algorithm expects:

  * no deadlock (the event queue never drains while an agent is blocked),
  * S_x(j) is overwritten only after all four warps of group x pulled S_x(j-1) into registers,
  * P_x(j) is written only after PV_x(j-1) has executed, PV_x(j) reads P_x(j) of all four warps,
  * O_x receives every key tile of the item exactly once and is read by the epilogue after the last PV,
  * K / V / Q smem stages are never refilled while an MMA that reads them is still pending.

mbarrier model: `count` arrivals complete a phase; wait(parity) succeeds when the barrier's current phase parity differs
from `parity` (PTX mbarrier.try_wait.parity).  tcgen05.commit arrives when all MMAs issued before it have executed
(in-order tensor pipe).  Run: python tools/protocol_sim.py [trials]
"""
from __future__ import annotations

import heapq
import random
import sys


class Barrier:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"{self.name}: more arrivals than the barrier expects"
        if self.pending == 0:
            self.pending = self.count
            self.phase += 1

    def test(self, parity):
        return (self.phase & 1) != parity


class Sim:
    def __init__(self, rng):
        self.rng, self.t, self.q, self.seq = rng, 0, [], 0
        self.agents = {}
        self.blocked = {}
        # in-order tensor pipe
        self.mma_free_at = 0

    def at(self, dt, fn):
        self.seq += 1
        heapq.heappush(self.q, (self.t + dt, self.seq, fn))

    def spawn(self, name, gen):
        self.agents[name] = gen
        self.at(0, lambda: self._step(name))

    def _step(self, name):
        gen = self.agents.get(name)
        if gen is None:
            return
        try:
            op = next(gen)
        except StopIteration:
            del self.agents[name]
            return
        if op[0] == "delay":
            self.at(op[1], lambda: self._step(name))
        elif op[0] == "wait":
            _, bar, parity = op
            if bar.test(parity):
                self.at(self.rng.randint(1, 20), lambda: self._step(name))
            else:
                self.blocked[name] = (bar, parity)
        else:
            raise ValueError(op)

    def poll(self):
        for name, (bar, parity) in list(self.blocked.items()):
            if bar.test(parity):
                del self.blocked[name]
                self.at(self.rng.randint(1, 30), lambda n=name: self._step(n))

    def mma(self, cycles, effect):
        """queue an op on the in-order tensor pipe; `effect` runs when it has executed"""
        start = max(self.t, self.mma_free_at)
        self.mma_free_at = start + cycles
        self.seq += 1
        heapq.heappush(self.q, (self.mma_free_at, self.seq, effect))

    def commit(self, bar):
        self.mma(0, bar.arrive)

    def run(self):
        while self.q:
            self.t, _, fn = heapq.heappop(self.q)
            fn()
            self.poll()
        assert not self.agents, f"DEADLOCK at t={self.t}: blocked " + ", ".join(
            f"{n} on {b.name} parity {p} (phase {b.phase})" for n, (b, p) in self.blocked.items())


def simulate_attn2q(rng, n_ctas_items, n_kv, stages=4, verbose=False, split=False, single_xchg_buffer=False):
    """one CTA processing `n_ctas_items` work items of `n_kv` key tiles each.  `split` = attn2q_split_kernel: two warps per lane
    quarter and query tile (64 key columns each), eight arrivals on s_free / p_ready / o_empty, the row maximum exchanged per key tile
    through a shared-memory slot per (tile, half, row) + a 64-thread named barrier, double-buffered by the tile parity;
    `single_xchg_buffer` drops the double buffering (negative control: a fast half overwrites the slot its partner has not read)."""
    sim = Sim(rng)
    S = stages
    B = lambda n, c: Barrier(n, c)
    q_full, q_empty = B("q_full", 1), B("q_empty", 1)
    k_full = [B(f"k_full{i}", 1) for i in range(S)]
    k_empty = [B(f"k_empty{i}", 1) for i in range(S)]
    v_full = [B(f"v_full{i}", 1) for i in range(S)]
    v_empty = [B(f"v_empty{i}", 1) for i in range(S)]
    s_full = [B(f"s_full{x}", 1) for x in range(2)]
    W = 8 if split else 4   # softmax warps per query tile
    s_free = [B(f"s_free{x}", W) for x in range(2)]
    p_ready = [B(f"p_ready{x}", W) for x in range(2)]
    pv_done = [B(f"pv_done{x}", 1) for x in range(2)]
    o_empty = [B(f"o_empty{x}", W) for x in range(2)]
    # split: named barrier of the two warps that share (tile x, quarter qd) — modelled as a 2-arrival barrier polled by phase — and the
    # exchange slots [parity][x][half][qd] holding the tile index whose half-maximum they carry
    pair_bar = [[B(f"pair{x}.{qd}", 2) for qd in range(4)] for x in range(2)]
    xchg = [[[[None] * 4 for _ in range(2)] for _ in range(2)] for _ in range(2)]

    # ---- data model
    q_smem = {"tag": None, "readers": 0}
    k_smem = [{"tag": None, "readers": 0} for _ in range(S)]
    v_smem = [{"tag": None, "readers": 0} for _ in range(S)]
    s_tmem = [{"tag": None, "loaded": [True] * W} for _ in range(2)]     # loaded[w]: warp w has the content in registers
    p_tmem = [{"tags": [None] * W, "consumed": True} for _ in range(2)]  # per-warp lane quarter (and column half)
    o_tmem = [{"item": None, "tiles": [], "read": [True] * W} for _ in range(2)]
    results = []

    def tma_fill(buf, tag, bar):
        def land():
            assert buf["readers"] == 0, f"TMA overwrote a stage that an MMA still reads (tag {buf['tag']} -> {tag})"
            buf["tag"] = tag
            bar.arrive()
        sim.at(rng.randint(200, 1500), land)

    def producer():
        ks = vs = 0
        kph = vph = 0
        for it in range(n_ctas_items):
            yield ("wait", q_empty, (it & 1) ^ 1)
            tma_fill(q_smem, ("q", it), q_full)
            for j in range(n_kv):
                yield ("wait", k_empty[ks], kph ^ 1)
                tma_fill(k_smem[ks], ("k", it, j), k_full[ks])
                ks += 1
                if ks == S:
                    ks, kph = 0, kph ^ 1
                yield ("wait", v_empty[vs], vph ^ 1)
                tma_fill(v_smem[vs], ("v", it, j), v_full[vs])
                vs += 1
                if vs == S:
                    vs, vph = 0, vph ^ 1
                yield ("delay", rng.randint(1, 40))

    def mma_warp():
        st = {"ks": 0, "vs": 0, "kph": 0, "vph": 0}
        g = 0

        def issue_s(gg, it, j, last_of_item):
            ks = st["ks"]
            yield ("wait", k_full[ks], st["kph"])
            for x in range(2):
                if gg > 0:
                    yield ("wait", s_free[x], (gg - 1) & 1)
                kb, qb = k_smem[ks], q_smem
                assert kb["tag"] == ("k", it, j), f"S({it},{j}) sees K stage tag {kb['tag']}"
                assert qb["tag"] == ("q", it), f"S({it},{j}) sees Q tag {qb['tag']}"
                kb["readers"] += 1
                qb["readers"] += 1

                def effect(x=x, kb=kb, qb=qb):
                    assert all(s_tmem[x]["loaded"]), f"S_{x}({it},{j}) overwrote scores that were not loaded yet"
                    s_tmem[x]["tag"] = (it, j)
                    s_tmem[x]["loaded"] = [False] * W
                    kb["readers"] -= 1
                    qb["readers"] -= 1
                sim.mma(rng.choice([200, 256, 300]), effect)
                sim.commit(s_full[x])
            sim.commit(k_empty[ks])
            if last_of_item:
                sim.commit(q_empty)
            st["ks"] += 1
            if st["ks"] == S:
                st["ks"], st["kph"] = 0, st["kph"] ^ 1

        for it in range(n_ctas_items):
            yield ("wait", q_full, it & 1)
            for j in range(n_kv):
                if j == 0:
                    yield from issue_s(g, it, 0, n_kv == 1)
                if j + 1 < n_kv:
                    yield from issue_s(g + 1, it, j + 1, j + 2 == n_kv)
                vs = st["vs"]
                yield ("wait", v_full[vs], st["vph"])
                for x in range(2):
                    yield ("wait", p_ready[x], g & 1)
                    if j == 0 and it > 0:
                        yield ("wait", o_empty[x], (it - 1) & 1)
                    vb = v_smem[vs]
                    assert vb["tag"] == ("v", it, j), f"PV({it},{j}) sees V stage tag {vb['tag']}"
                    vb["readers"] += 1

                    def effect(x=x, vb=vb, it=it, j=j):
                        assert p_tmem[x]["tags"] == [(it, j)] * W, f"PV_{x}({it},{j}) read P tags {p_tmem[x]['tags']}"
                        p_tmem[x]["consumed"] = True
                        o = o_tmem[x]
                        if j == 0:
                            assert all(o["read"]), f"PV_{x}({it},0) overwrote an O tile the epilogue had not read"
                            o["item"], o["tiles"], o["read"] = it, [], [False] * W
                        assert o["item"] == it
                        o["tiles"].append(j)
                        vb["readers"] -= 1
                    sim.mma(rng.choice([100, 128, 160]), effect)
                    sim.commit(pv_done[x])
                sim.commit(v_empty[vs])
                st["vs"] += 1
                if st["vs"] == S:
                    st["vs"], st["vph"] = 0, st["vph"] ^ 1
                g += 1
                yield ("delay", rng.randint(1, 30))

    def softmax_warp(x, w):
        n = 0
        for it in range(n_ctas_items):
            for j in range(n_kv):
                yield ("wait", s_full[x], n & 1)
                yield ("delay", rng.randint(50, 400))          # tcgen05.ld + wait
                assert s_tmem[x]["tag"] == (it, j), f"group {x} warp {w} loaded S tag {s_tmem[x]['tag']}, wants {(it, j)}"
                s_tmem[x]["loaded"][w] = True
                s_free[x].arrive()
                yield ("delay", rng.randint(50, 300))          # mask + row max
                if split:  # exchange the half-row maximum with the partner warp (same tile, same quarter, other column half)
                    qd, hf = w & 3, w >> 2
                    par = 0 if single_xchg_buffer else n & 1
                    xchg[par][x][hf][qd] = n
                    bar = pair_bar[x][qd]
                    ph = bar.phase
                    bar.arrive()
                    yield ("wait", bar, ph & 1)             # bar.sync 64: returns once both warps arrived
                    yield ("delay", rng.choice([5, 50, 4000]))   # the read may be arbitrarily late: only the barriers order it
                    assert xchg[par][x][hf ^ 1][qd] == n, \
                        f"tile {x} quarter {qd}: half {hf} read the partner's maximum of key tile {xchg[par][x][hf ^ 1][qd]}, wants {n}"
                if n > 0:
                    yield ("wait", pv_done[x], (n - 1) & 1)
                if j > 0 and rng.random() < 0.3:                # rescale O_x (needs PV(j-1) done: asserted here)
                    assert o_tmem[x]["tiles"] == list(range(j)), f"rescale of O_{x} at tile {j} sees {o_tmem[x]['tiles']}"
                    yield ("delay", rng.randint(50, 200))
                # P write: the previous P of this quarter must have been consumed by PV(n-1)
                if n > 0:
                    assert p_tmem[x]["consumed"], f"group {x} warp {w} overwrote P before PV consumed it"
                yield ("delay", rng.randint(300, 1500))         # exponentials + tcgen05.st
                p_tmem[x]["tags"][w] = (it, j)
                if all(t == (it, j) for t in p_tmem[x]["tags"]):
                    p_tmem[x]["consumed"] = False
                p_ready[x].arrive()
                n += 1
            yield ("wait", pv_done[x], (n - 1) & 1)
            o = o_tmem[x]
            assert o["item"] == it and o["tiles"] == list(range(n_kv)), f"epilogue of item {it} sees O = {o['item']}, {o['tiles']}"
            yield ("delay", rng.randint(50, 400))
            o["read"][w] = True
            results.append((it, x, w))
            o_empty[x].arrive()

    sim.spawn("producer", producer())
    sim.spawn("mma", mma_warp())
    for x in range(2):
        for w in range(W):
            sim.spawn(f"softmax{x}.{w}", softmax_warp(x, w))
    sim.run()
    assert len(results) == n_ctas_items * 2 * W
    return sim.t


class Chan:
    """blocking mailbox (the running-max / row-sum hand-over between the two threads of a query row)"""

    def __init__(self, name):
        self.name, self.items, self.phase = name, [], 0

    def put(self, v):
        self.items.append(v)

    def test(self, _parity):
        return bool(self.items)

    def get(self):
        return self.items.pop(0)


def simulate_tfused(rng, n_items, num_kb, stages=4):
    """csrc/attention_tfused_tcgen05.cu: projection ring -> [Q K V] accumulators -> convert (Q16 in TMEM, K / V tiles in smem)
    -> S -> softmax -> PV -> epilogue, with the NEXT item's projection issued under the current item's softmax"""
    sim = Sim(rng)
    S = stages
    full = [Barrier(f"full{i}", 1) for i in range(S)]
    empty = [Barrier(f"empty{i}", 1) for i in range(S)]
    qkv_full, conv_done, s_full, p_ready, o_full = Barrier("qkv_full", 1), Barrier("conv_done", 4), Barrier("s_full", 1), Barrier("p_ready", 4), Barrier("o_full", 1)
    ring = [{"tag": None, "readers": 0} for _ in range(S)]
    acc = {"tag": None, "conv": [True] * 4}           # fp32 [Q K V]: which item, which warps have converted it
    q16 = {"tags": [None] * 4}
    kv = {"tags": [None] * 4, "readers": 0}
    sp = {"tag": None, "p": [None] * 4}
    o = {"tag": None, "read": [True] * 4}
    done = []

    def producer():
        st, ph = 0, 0
        for it in range(n_items):
            for kb in range(num_kb):
                yield ("wait", empty[st], ph ^ 1)
                buf, tag, bar = ring[st], (it, kb), full[st]

                def land(buf=buf, tag=tag, bar=bar):
                    assert buf["readers"] == 0, f"TMA overwrote ring stage {buf['tag']} -> {tag}"
                    buf["tag"] = tag
                    bar.arrive()
                sim.at(rng.randint(200, 1500), land)
                st += 1
                if st == S:
                    st, ph = 0, ph ^ 1

    def mma():
        state = {"st": 0, "ph": 0, "next": 0}

        def issue_qkv():
            it = state["next"]
            state["next"] += 1
            for kb in range(num_kb):
                st = state["st"]
                yield ("wait", full[st], state["ph"])
                buf = ring[st]
                assert buf["tag"] == (it, kb), (buf["tag"], it, kb)
                buf["readers"] += 1

                def effect(buf=buf, it=it, kb=kb):
                    if kb == 0:
                        assert all(acc["conv"]), f"projection of item {it} overwrote accumulators that were not converted yet"
                        acc["tag"], acc["conv"] = it, [False] * 4
                    assert acc["tag"] == it
                    buf["readers"] -= 1
                sim.mma(rng.choice([60, 96, 130]), effect)
                sim.commit(empty[st])
                state["st"] += 1
                if state["st"] == S:
                    state["st"], state["ph"] = 0, state["ph"] ^ 1
            sim.commit(qkv_full)

        if n_items > 0:
            yield from issue_qkv()
        for it in range(n_items):
            yield ("wait", conv_done, it & 1)
            kv["readers"] += 1

            def s_effect(it=it):
                assert q16["tags"] == [it] * 4 and kv["tags"] == [it] * 4, (it, q16["tags"], kv["tags"])
                sp["tag"], sp["p"] = it, [None] * 4
                kv["readers"] -= 1
            sim.mma(rng.choice([200, 256]), s_effect)
            sim.commit(s_full)
            if it + 1 < n_items:
                yield from issue_qkv()
            yield ("wait", p_ready, it & 1)
            kv["readers"] += 1

            def pv_effect(it=it):
                assert sp["tag"] == it and sp["p"] == [it] * 4 and kv["tags"] == [it] * 4
                assert all(o["read"]), "PV overwrote an O tile the epilogue had not read"
                o["tag"], o["read"] = it, [False] * 4
                kv["readers"] -= 1
            sim.mma(rng.choice([200, 256]), pv_effect)
            sim.commit(o_full)

    def compute(w):
        for it in range(n_items):
            yield ("wait", qkv_full, it & 1)
            yield ("delay", rng.randint(100, 600))
            assert acc["tag"] == it, f"warp {w} converts accumulators of item {acc['tag']}, wants {it}"
            assert kv["readers"] == 0, "K / V tiles rewritten while an MMA still reads them"
            q16["tags"][w] = it
            kv["tags"][w] = it
            acc["conv"][w] = True
            conv_done.arrive()
            yield ("wait", s_full, it & 1)
            yield ("delay", rng.randint(300, 1500))
            assert sp["tag"] == it
            sp["p"][w] = it
            p_ready.arrive()
            yield ("wait", o_full, it & 1)
            assert o["tag"] == it
            yield ("delay", rng.randint(50, 400))
            o["read"][w] = True
            done.append((it, w))

    sim.spawn("producer", producer())
    sim.spawn("mma", mma())
    for w in range(4):
        sim.spawn(f"compute{w}", compute(w))
    sim.run()
    assert len(done) == 4 * n_items
    return sim.t


def simulate_tfused2(rng, n_items, num_kb, stages=4, wrong_order=False):
    """tattn_fused2_kernel: two convert / softmax warpgroups on alternate items, one TMEM slot each.  A slot's 192 accumulator columns
    are RE-USED IN PLACE — Q fp16 over the Q accumulator (same thread, after reading it), S over the K / V accumulators once all four
    warps converted them, P over S — so the model tracks per slot: which item's accumulators / Q16 / S / P the columns hold, and that
      * QKV(i+2) never lands on a slot whose S(i) / P(i) is still to be read by PV(i) or whose Q16(i) by S(i) (in-order pipe + issue order),
      * S(i) is issued only after all four warps of the group converted item i (it overwrites the K / V accumulators),
      * the slot's K / V tiles are rewritten (convert of item i+2) only after S(i) and PV(i) executed,
      * O of the slot is overwritten by PV(i+2) only after the group's epilogue read O(i).
    MMA order:  QKV(0) QKV(1) | S(i) PV(i-1) QKV(i+1) | ... | PV(n-1).  `wrong_order` issues QKV(i+1) BEFORE PV(i-1) (the slot is
    still in use): the model must reject it — that is the negative control of the test."""
    sim = Sim(rng)
    S = stages
    full = [Barrier(f"full{i}", 1) for i in range(S)]
    empty = [Barrier(f"empty{i}", 1) for i in range(S)]
    qkv_full = [Barrier(f"qkv_full{g}", 1) for g in range(2)]
    conv_done = [Barrier(f"conv_done{g}", 4) for g in range(2)]
    s_full = [Barrier(f"s_full{g}", 1) for g in range(2)]
    p_ready = [Barrier(f"p_ready{g}", 4) for g in range(2)]
    o_full = [Barrier(f"o_full{g}", 1) for g in range(2)]
    ring = [{"tag": None, "readers": 0} for _ in range(S)]
    # slot state: acc = item whose fp32 [Q K V] the columns hold (None once S overwrote K / V), conv = warps that converted it,
    # q16 / p per warp, s = item whose scores are in the columns, pending = MMAs issued that still read Q16 / P of the slot
    slot = [{"acc": None, "conv": [True] * 4, "q16": [None] * 4, "s": None, "p": [None] * 4, "pending": 0,
             "kv": [None] * 4, "kv_readers": 0, "o": None, "o_read": [True] * 4} for _ in range(2)]
    done = []

    def producer():
        st, ph = 0, 0
        for it in range(n_items):
            for kb in range(num_kb):
                yield ("wait", empty[st], ph ^ 1)
                buf, tag, bar = ring[st], (it, kb), full[st]

                def land(buf=buf, tag=tag, bar=bar):
                    assert buf["readers"] == 0, f"TMA overwrote ring stage {buf['tag']} -> {tag}"
                    buf["tag"] = tag
                    bar.arrive()
                sim.at(rng.randint(200, 1500), land)
                st += 1
                if st == S:
                    st, ph = 0, ph ^ 1

    def mma():
        state = {"st": 0, "ph": 0}

        def issue_qkv(it):
            sl = slot[it & 1]
            for kb in range(num_kb):
                st = state["st"]
                yield ("wait", full[st], state["ph"])
                buf = ring[st]
                assert buf["tag"] == (it, kb), (buf["tag"], it, kb)
                buf["readers"] += 1

                def effect(buf=buf, it=it, kb=kb, sl=sl):
                    if kb == 0:
                        assert sl["pending"] == 0, f"projection of item {it} landed on a slot whose Q16 / P an MMA still reads"
                        assert all(sl["conv"]), f"projection of item {it} overwrote accumulators that were not converted yet"
                        assert it < 2 or (sl["s"] == it - 2 and sl["p"] == [it - 2] * 4), "slot re-used before its previous item's PV"
                        sl["acc"], sl["conv"], sl["q16"], sl["s"], sl["p"] = it, [False] * 4, [None] * 4, None, [None] * 4
                    assert sl["acc"] == it
                    buf["readers"] -= 1
                sim.mma(rng.choice([60, 96, 130]), effect)
                sim.commit(empty[st])
                state["st"] += 1
                if state["st"] == S:
                    state["st"], state["ph"] = 0, state["ph"] ^ 1
            sim.commit(qkv_full[it & 1])

        def issue_pv(it):
            g = it & 1
            sl = slot[g]
            yield ("wait", p_ready[g], (it >> 1) & 1)
            sl["kv_readers"] += 1
            sl["pending"] += 1

            def pv_effect(it=it, sl=sl):
                assert sl["s"] == it and sl["p"] == [it] * 4 and sl["kv"] == [it] * 4
                assert all(sl["o_read"]), "PV overwrote an O tile the epilogue had not read"
                sl["o"], sl["o_read"] = it, [False] * 4
                sl["kv_readers"] -= 1
                sl["pending"] -= 1
            sim.mma(rng.choice([200, 256]), pv_effect)
            sim.commit(o_full[g])

        if n_items > 0:
            yield from issue_qkv(0)
        if n_items > 1:
            yield from issue_qkv(1)
        for it in range(n_items):
            g = it & 1
            sl = slot[g]
            yield ("wait", conv_done[g], (it >> 1) & 1)
            sl["kv_readers"] += 1
            sl["pending"] += 1

            def s_effect(it=it, sl=sl):
                assert sl["q16"] == [it] * 4 and sl["kv"] == [it] * 4 and all(sl["conv"]), (it, sl["q16"], sl["kv"], sl["conv"])
                sl["acc"], sl["s"], sl["p"] = None, it, [None] * 4   # S lands on the K / V accumulators
                sl["kv_readers"] -= 1
                sl["pending"] -= 1
            sim.mma(rng.choice([200, 256]), s_effect)
            sim.commit(s_full[g])
            if it >= 1:
                if wrong_order and it + 1 < n_items:
                    yield from issue_qkv(it + 1)
                yield from issue_pv(it - 1)
                if not wrong_order and it + 1 < n_items:
                    yield from issue_qkv(it + 1)
        if n_items > 0:
            yield from issue_pv(n_items - 1)

    def compute(g, w):
        sl = slot[g]
        k = 0
        for it in range(g, n_items, 2):
            yield ("wait", qkv_full[g], k & 1)
            yield ("delay", rng.randint(100, 600))
            assert sl["acc"] == it, f"group {g} warp {w} converts accumulators of item {sl['acc']}, wants {it}"
            assert sl["kv_readers"] == 0, "K / V tiles rewritten while an MMA still reads them"
            sl["q16"][w] = it   # over this lane's Q accumulator, after reading it
            sl["kv"][w] = it
            sl["conv"][w] = True
            conv_done[g].arrive()
            yield ("wait", s_full[g], k & 1)
            yield ("delay", rng.randint(300, 1500))
            assert sl["s"] == it
            sl["p"][w] = it
            p_ready[g].arrive()
            yield ("wait", o_full[g], k & 1)
            assert sl["o"] == it
            yield ("delay", rng.randint(50, 400))
            sl["o_read"][w] = True
            done.append((it, w))
            k += 1

    sim.spawn("producer", producer())
    sim.spawn("mma", mma())
    for g in range(2):
        for w in range(4):
            sim.spawn(f"compute{g}{w}", compute(g, w))
    sim.run()
    assert len(done) == 4 * n_items
    return sim.t


def main(trials=300):
    rng = random.Random(1234)
    worst = 0
    for trial in range(trials):
        items = rng.choice([1, 2, 3, 5])
        n_kv = rng.choice([1, 2, 3, 4, 7, 8, 32])
        stages = rng.choice([2, 3, 4])
        worst = max(worst, simulate_attn2q(random.Random(rng.getrandbits(32)), items, n_kv, stages))
    print(f"attn2q protocol: {trials} randomised schedules, no deadlock, no buffer hazard (longest run {worst} cycles)")
    worst = 0
    for trial in range(max(trials // 4, 1)):
        worst = max(worst, simulate_attn2q(random.Random(rng.getrandbits(32)), rng.choice([1, 2, 3]), rng.choice([1, 2, 16, 17, 32]),
                                           rng.choice([2, 3, 4]), split=True))
    print(f"attn2q split (two threads per row) protocol: {max(trials // 4, 1)} randomised schedules, no deadlock, no hazard (longest run {worst} cycles)")
    worst = 0
    for trial in range(trials):
        worst = max(worst, simulate_tfused(random.Random(rng.getrandbits(32)), rng.choice([1, 2, 3, 6]), rng.choice([1, 5, 8, 10, 20]),
                                           rng.choice([2, 3, 4])))
    print(f"fused temporal attention protocol: {trials} randomised schedules, no deadlock, no buffer hazard (longest run {worst} cycles)")
    worst = 0
    for trial in range(trials):
        worst = max(worst, simulate_tfused2(random.Random(rng.getrandbits(32)), rng.choice([1, 2, 3, 4, 7, 10]), rng.choice([1, 5, 8, 10]),
                                            rng.choice([2, 3, 4])))
    print(f"two-slot fused temporal attention protocol: {trials} randomised schedules, no deadlock, no aliasing hazard (longest run {worst} cycles)")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 300)
