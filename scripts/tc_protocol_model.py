"""Discrete-event model of the barrier protocol of the two-chunk tensor-core kernel (rlaopt_b200/csrc/kmm_tc.cu, DUAL
instantiations with the sliced epilogue), used to look for orderings that break a data invariant.

What is modelled (one CTA, no pairs):
  * mbarriers with the hardware's phase / parity semantics (a parity wait is ambiguous by two phases -- aliasing is part of
    the model), arrival counts as in the kernel;
  * the producer thread (A ring of SA column-tile images, V ring of SV records, strictly sequential issue, every bulk copy
    with its own random latency and occasional latency spikes);
  * the two MMA1 issue warps (even / odd sub-tiles), the MMA2 issue warp (two chunks per sub-tile), one in-order tensor
    pipe, tcgen05.commit = "arrive when everything this warp issued before has completed";
  * the eight epilogue warps (two warpgroups) walking all sub-tiles with the slot schedules of the kernel:
      "late"   quarter per slot, P'(u) announced at the end of slot B(u-1)           (shipped, RLAOPT_B200_TC_DUAL_OVERLAP=5)
      "early"  row extreme + q0 | q1 + q2 | q3 + announce | drain only                (the schedule that mis-computes)
      "defer"  as "early", announcement deferred to the start of slot B(u-1)
      "last"   dual_mode 1 (drains behind the pointwise stage; the default two-chunk kernels, 64 < d <= 128: NB 2, SA 2, SV 4)
  * the contents of every A stage, V stage, S/P buffer and O buffer, checked at every read and write:
      MMA1 reads the image of ITS tile, fully landed;  the pointwise stage reads a complete S of its tile;  MMA2 reads a
      complete P' of its tile and the V record of its chunk;  a drain reads the complete O of its tile and chunk;  nothing
      is overwritten while a reader is still due.

    python scripts/tc_protocol_model.py [schedule] [seeds] [tiles]

The model assumes an in-order tensor pipe; `--ooo` lets operations of different issue warps overlap in time (each warp's
own operations stay ordered).  `--limit=L` samples the barriers of a multi-barrier poll one after the other, each
try_wait suspending up to L cycles (the PTX semantics; the idealised default samples them at one instant) and runs both
orders of the MMA1 issue warps' wait: with the a_full sample first, L >= ~6000 lets MMA1 read an A stage whose bulk copy
is in flight; with the p_free guard first (the kernel's order) no schedule violates anything.
"""
from __future__ import annotations

import heapq
import random
import sys

NB, SA, SV = 3, 3, 4
MMA1_CYC, MMA2_CYC = 384, 768


class Violation(Exception):
    pass


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def test(self, parity):  # mbarrier.try_wait.parity
        return (self.phase & 1) != parity


class Sim:
    def __init__(self, T, schedule, seed, ooo=False, spike=0.02, try_wait_limit=0, guard_first=True, nb=NB, sa=SA, sv=SV,
                 exact_guard=False):
        # try_wait_limit = 0: the barriers of one poll are sampled at the same instant (idealised); > 0: sampled one after
        # the other, each mbarrier.try_wait suspending the thread until its phase completes or the limit (cycles) expires
        # -- a sample taken early in the poll can be stale when the last one returns.  guard_first: order of the two
        # barriers in the MMA1 issue warps' wait (True = p_free then a_full, the kernel since the fix; False = before).
        self.T, self.schedule, self.ooo = T, schedule, ooo
        self.NB, self.SA, self.SV = nb, sa, sv
        self.exact_guard = exact_guard
        self.L, self.guard_first = try_wait_limit, guard_first
        self.single = []
        self.rng = random.Random(seed)
        self.spike = spike
        self.now = 0
        self.events = []  # (time, seq, fn)
        self.seq = 0
        self.waiters = []  # (agent generator, condition)
        b = lambda c: Bar(c)
        self.a_full = [b(1) for _ in range(self.SA)]
        self.a_empty = [b(1) for _ in range(self.SA)]
        self.v_full = [b(1) for _ in range(self.SV)]
        self.v_empty = [b(1) for _ in range(self.SV)]
        self.s_full = [b(1) for _ in range(self.NB)]
        self.p_full = [b(4) for _ in range(self.NB)]
        self.p_free = [b(1) for _ in range(self.NB)]
        self.o_full = [b(1) for _ in range(2)]
        self.o_free = [b(8) for _ in range(2)]
        # contents
        self.a_stage = [None] * self.SA        # tile id, or ("loading", tile)
        self.v_stage = [None] * self.SV        # (tile, chunk) or ("loading", ...)
        self.sp = [None] * self.NB             # ("S", t, done) / ("P", t, quarters_done_by_warp dict)
        self.sp_busy = [None] * self.NB        # tensor op currently writing / reading: (kind, t, end)
        self.o = [None] * 2               # ("O", t, done)
        self.o_readers = [0, 0]           # drains in progress
        self.pipe_end = 0                 # in-order tensor pipe
        self.warp_end = {}                # per issue warp: completion time of its last op
        self.log = []

    # ---- event machinery ------------------------------------------------------------------------------------------
    def at(self, t, fn):
        self.seq += 1
        heapq.heappush(self.events, (t, self.seq, fn))

    def spawn(self, gen):
        self.step(gen)

    def step(self, gen):
        try:
            req = next(gen)
        except StopIteration:
            return
        kind = req[0]
        if kind == "delay":
            self.at(self.now + req[1], lambda g=gen: self.step(g))
        elif kind == "wait":  # list of (bar, parity), all true in the same poll
            if self.L > 0 and len(req[1]) > 1:
                self.seq_poll(gen, req[1])
            else:
                self.waiters.append((gen, req[1]))
                self.poll()
        else:
            raise ValueError(kind)

    def seq_poll(self, gen, conds):
        """One pass of `try_wait c0; try_wait c1; ...; and` -- repeated until every sample was true."""
        results = []

        def next_cond():
            i = len(results)
            if i == len(conds):
                if all(results):
                    self.at(self.now + self.rng.randint(20, 120), lambda: self.step(gen))
                else:
                    self.at(self.now + 4, lambda: self.seq_poll(gen, conds))
                return
            bar, par = conds[i]
            if bar.test(par):
                results.append(True)
                self.at(self.now + 2, next_cond)
                return
            w = {"bar": bar, "par": par, "done": False}

            def finish(val):
                if w["done"]:
                    return
                w["done"] = True
                results.append(val)
                self.at(self.now + 2, next_cond)
            w["cb"] = finish
            self.single.append(w)
            self.at(self.now + self.L, lambda: finish(False))
        next_cond()

    def poll(self):
        for w in list(self.single):
            if w["done"]:
                self.single.remove(w)
            elif w["bar"].test(w["par"]):
                self.single.remove(w)
                w["cb"](True)
        again = True
        while again:
            again = False
            for i, (gen, conds) in enumerate(self.waiters):
                if all(bar.test(p) for bar, p in conds):
                    self.waiters.pop(i)
                    # a successful poll is seen 20-120 cycles later
                    self.at(self.now + self.rng.randint(20, 120), lambda g=gen: self.step(g))
                    again = True
                    break

    def arrive(self, bar):
        bar.arrive()
        self.poll()

    def run(self):
        while self.events:
            t, _, fn = heapq.heappop(self.events)
            self.now = t
            fn()
        if self.waiters or any(not w["done"] for w in self.single):
            raise Violation(f"deadlock: {len(self.waiters)} agents waiting at the end")

    def fail(self, msg):
        raise Violation(f"t={self.now}: {msg}")

    def load_latency(self):
        lat = self.rng.randint(700, 1500)
        if self.rng.random() < self.spike:
            lat += self.rng.randint(3000, 40000)
        return lat

    # ---- tensor pipe ---------------------------------------------------------------------------------------------------
    def issue(self, warp, dur, on_start, on_end, commits):
        """Queue one group of MMAs; commits fire when everything `warp` issued so far has completed."""
        if self.ooo:
            start = max(self.now, self.warp_end.get(warp, 0))
        else:
            start = max(self.now, self.pipe_end)
        end = start + dur
        self.pipe_end = max(self.pipe_end, end)
        self.warp_end[warp] = max(self.warp_end.get(warp, 0), end)
        self.at(start, on_start)
        self.at(end, on_end)
        fire = self.warp_end[warp]
        for bar in commits:
            self.at(fire, lambda b=bar: self.arrive(b))

    # ---- agents ----------------------------------------------------------------------------------------------------------
    def producer(self):
        sa = sv = 0
        pha = phv = 1
        for u in range(self.T):
            yield ("wait", [(self.a_empty[sa], pha)])
            if isinstance(self.a_stage[sa], tuple):
                self.fail(f"producer overwrites A stage {sa} while {self.a_stage[sa]}")
            self.a_stage[sa] = ("loading", u)

            def landed(s=sa, t=u):
                self.a_stage[s] = t
                self.arrive(self.a_full[s])
            self.at(self.now + self.load_latency(), landed)
            for c in range(2):
                yield ("wait", [(self.v_empty[sv], phv)])
                self.v_stage[sv] = ("loading", u, c)

                def vlanded(s=sv, t=u, cc=c):
                    self.v_stage[s] = (t, cc)
                    self.arrive(self.v_full[s])
                self.at(self.now + self.load_latency(), vlanded)
                sv += 1
                if sv == self.SV:
                    sv, phv = 0, phv ^ 1
                yield ("delay", 10)
            sa += 1
            if sa == self.SA:
                sa, pha = 0, pha ^ 1
            yield ("delay", 10)

    def mma1(self, par):
        b1, sa = par % self.NB, par % self.SA
        use1, pha = (par // self.NB) & 1, (par // self.SA) & 1
        for t1 in range(par, self.T, 2):
            conds = [(self.p_free[b1], use1 ^ 1), (self.a_full[sa], pha)]
            conds = conds if self.guard_first else conds[::-1]
            if self.exact_guard:  # MMA1 of the stage's previous tile has completed (the barrier the producer waits on too)
                conds.insert(0, (self.a_empty[sa], pha ^ 1))
            yield ("wait", conds)

            def start(s=sa, b=b1, t=t1):
                if self.a_stage[s] != t:
                    self.fail(f"MMA1({t}) reads A stage {s} holding {self.a_stage[s]}")
                cur = self.sp[b]
                if cur is not None and not (cur[0] == "P" and cur[1] == t - self.NB and cur[2] == "consumed"):
                    self.fail(f"MMA1({t}) overwrites S/P buffer {b} holding {cur}")
                self.sp[b] = ("S", t, False)

            def end(s=sa, b=b1, t=t1):
                if self.a_stage[s] != t:
                    self.fail(f"A stage {s} changed under MMA1({t}): {self.a_stage[s]}")
                self.sp[b] = ("S", t, True)
            self.issue(("mma1", par), MMA1_CYC, start, end, [self.s_full[b1], self.a_empty[sa]])
            yield ("delay", 60)
            b1 += 2
            if b1 >= self.NB:
                b1, use1 = b1 - self.NB, use1 ^ 1
            sa += 2
            if sa >= self.SA:
                sa, pha = sa - self.SA, pha ^ 1

    def mma2(self):
        b2 = sv = 0
        use2 = phv = 0
        for u in range(self.T):
            opar = u & 1
            for c in range(2):
                conds = [(self.v_full[sv], phv), (self.o_free[c], opar ^ 1)]
                if c == 0:
                    conds.insert(0, (self.p_full[b2], use2))  # the kernel's order: p_full, v_full, o_free
                yield ("wait", conds)

                def start(b=b2, s=sv, t=u, cc=c):
                    cur = self.sp[b]
                    if not (cur and cur[0] == "P" and cur[1] == t and cur[2] == "complete"):
                        self.fail(f"MMA2({t},{cc}) reads P' buffer {b} holding {cur}")
                    if self.v_stage[s] != (t, cc):
                        self.fail(f"MMA2({t},{cc}) reads V stage {s} holding {self.v_stage[s]}")
                    if self.o_readers[cc]:
                        self.fail(f"MMA2({t},{cc}) overwrites O[{cc}] under {self.o_readers[cc]} drains")
                    if self.o[cc] is not None and self.o[cc][3] != 8:
                        self.fail(f"MMA2({t},{cc}) overwrites O[{cc}] = {self.o[cc]} before all warps drained it")
                    self.o[cc] = ("O", t, False, 0)

                def end(b=b2, t=u, cc=c):
                    self.o[cc] = ("O", t, True, 0)
                    if cc == 1:
                        self.sp[b] = ("P", t, "consumed")
                commits = [self.v_empty[sv]] + ([self.p_free[b2]] if c == 1 else []) + [self.o_full[c]]
                self.issue("mma2", MMA2_CYC, start, end, commits)
                yield ("delay", 60)
                sv += 1
                if sv == self.SV:
                    sv, phv = 0, phv ^ 1
            b2 += 1
            if b2 == self.NB:
                b2, use2 = 0, use2 ^ 1

    def epi_warp(self, g, q):
        T, sched = self.T, self.schedule
        speed = 1.0 + 0.05 * q + 0.02 * self.rng.random()
        pending = [None]

        def drain(t, ch):
            yield ("wait", [(self.o_full[ch], t & 1)])
            cur = self.o[ch]
            if not (cur and cur[1] == t and cur[2]):
                self.fail(f"warp {g}.{q} drains O[{ch}] for tile {t}, holds {cur}")
            self.o_readers[ch] += 1
            yield ("delay", int(300 * speed))
            if self.o[ch][1] != t:
                self.fail(f"O[{ch}] overwritten under the drain of tile {t} by warp {g}.{q}: {self.o[ch]}")
            self.o_readers[ch] -= 1
            self.o[ch] = self.o[ch][:3] + (self.o[ch][3] + 1,)
            self.arrive(self.o_free[ch])

        def check_s(u, what):
            cur = self.sp[u % self.NB]
            if not cur or cur[1] != u or (cur[0] == "S" and not cur[2]):
                self.fail(f"warp {g}.{q} {what} of tile {u}: buffer {u % self.NB} holds {cur}")

        def slice0(u):
            yield ("wait", [(self.s_full[u % self.NB], (u // self.NB) & 1)])
            check_s(u, "reads S")
            yield ("delay", int(650 * speed))
            check_s(u, "writes q0")
            self.note_quarter(u, g, q, 0)

        def quarter(u, sl):
            check_s(u, f"reads S quarter {sl}")
            yield ("delay", int(420 * speed))
            check_s(u, f"writes quarter {sl}")
            self.note_quarter(u, g, q, sl)

        def announce(u):
            self.arrive(self.p_full[u % self.NB])

        def publish(u, defer):
            if defer:
                pending[0] = u
            else:
                announce(u)

        if sched == "last":
            # dual_mode 1 (the default for 64 < d <= 128): a warpgroup runs the whole pointwise stage of its own sub-tile,
            # announces P', then drains both chunks of every sub-tile up to u - 1 (half of the columns each)
            nd = 0
            for u in list(range(g, T, 2)) + [T + g]:
                if u < T:
                    sva, svb = (2 * u) % self.SV, (2 * u + 1) % self.SV
                    yield ("wait", [(self.v_full[sva], ((2 * u) // self.SV) & 1), (self.v_full[svb], ((2 * u + 1) // self.SV) & 1),
                                    (self.s_full[u % self.NB], (u // self.NB) & 1)])
                    for ss, cc in ((sva, 0), (svb, 1)):
                        if self.v_stage[ss] != (u, cc):
                            self.fail(f"warp {g}.{q} reads the norms / scale of tile {u} from V stage {ss} holding {self.v_stage[ss]}")
                    check_s(u, "reads S")
                    yield ("delay", int(1800 * speed))
                    for sl in range(4):
                        self.note_quarter(u, g, q, sl)
                    announce(u)
                while nd <= min(u - 1, T - 1):
                    yield from drain(nd, 0)
                    yield from drain(nd, 1)
                    nd += 1
            return
        for t in range(-2, T):
            mine = (t & 1) == g
            if sched == "late":
                items = [("S0",), ("Q", 1)] if mine else [("Q", 2), ("Q", 3, "pub")]
            else:
                items = [("S0",), ("Q", 1, 2)] if mine else [("Q", 3, "pub"), ()]
            u = t + 2 if mine else t + 1
            for ch in range(2):
                if pending[0] is not None:
                    if t >= 0:
                        yield ("wait", [(self.o_full[ch], t & 1)])
                    announce(pending[0])
                    pending[0] = None
                if t >= 0:
                    yield from drain(t, ch)
                it = items[ch]
                if not it or not (0 <= u < T):
                    continue
                if it[0] == "S0":
                    yield from slice0(u)
                else:
                    for sl in it[1:]:
                        if sl == "pub":
                            publish(u, sched == "defer")
                        else:
                            yield from quarter(u, sl)

    def note_quarter(self, u, g, q, sl):
        b = u % self.NB
        cur = self.sp[b]
        if cur[0] == "S":
            cur = ("P", u, {})
        done = cur[2] if isinstance(cur[2], dict) else {}
        done[(q, sl)] = True
        self.sp[b] = ("P", u, "complete" if len(done) == 16 else done)



class SimSingle(Sim):
    """The one-chunk kernels (KP <= 64, two epilogue warpgroups that own alternate sub-tiles): per own tile a warpgroup waits
    for {V record (column norms, V scale), S}, writes P', announces it and drains the O buffer of its previous tile; MMA2
    contracts one chunk per tile into O[u % 2].  Default rings: the C2 plan (d = 128, k = 64): NB = 4, SA = 3, SV = 5 -- an
    odd V ring is shared between the two warpgroups, so a warpgroup's v_full parity wait is ambiguous while the OTHER
    warpgroup's record of tile u - SV is in flight, and the s_full wait of the same poll is its guard.
    epi_guard_first: s_full polled before v_full (False = the kernel's order, v_full first)."""

    def __init__(self, T, seed, epi_guard_first=False, nb=4, sa=3, sv=5, **kw):
        super().__init__(T, "single", seed, nb=nb, sa=sa, sv=sv, **kw)
        self.epi_guard_first = epi_guard_first
        self.o_free = [Bar(4) for _ in range(2)]

    def producer(self):
        sa = sv = 0
        pha = phv = 1
        for u in range(self.T):
            yield ("wait", [(self.a_empty[sa], pha)])
            self.a_stage[sa] = ("loading", u)

            def landed(s=sa, t=u):
                self.a_stage[s] = t
                self.arrive(self.a_full[s])
            self.at(self.now + self.load_latency(), landed)
            yield ("wait", [(self.v_empty[sv], phv)])
            self.v_stage[sv] = ("loading", u)

            def vlanded(s=sv, t=u):
                self.v_stage[s] = (t, 0)
                self.arrive(self.v_full[s])
            self.at(self.now + self.load_latency(), vlanded)
            sa += 1
            if sa == self.SA:
                sa, pha = 0, pha ^ 1
            sv += 1
            if sv == self.SV:
                sv, phv = 0, phv ^ 1
            yield ("delay", 10)

    def mma2(self):
        b2 = sv = 0
        use2 = phv = 0
        for u in range(self.T):
            ob, opar = u & 1, (u >> 1) & 1
            yield ("wait", [(self.p_full[b2], use2), (self.v_full[sv], phv), (self.o_free[ob], opar ^ 1)])

            def start(b=b2, s=sv, t=u, o=ob):
                cur = self.sp[b]
                if not (cur and cur[0] == "P" and cur[1] == t and cur[2] == "complete"):
                    self.fail(f"MMA2({t}) reads P' buffer {b} holding {cur}")
                if self.v_stage[s] != (t, 0):
                    self.fail(f"MMA2({t}) reads V stage {s} holding {self.v_stage[s]}")
                if self.o[o] is not None and self.o[o][3] != 4:
                    self.fail(f"MMA2({t}) overwrites O[{o}] = {self.o[o]} before it was drained")
                self.o[o] = ("O", t, False, 0)

            def end(b=b2, t=u, o=ob):
                self.o[o] = ("O", t, True, 0)
                self.sp[b] = ("P", t, "consumed")
            self.issue("mma2", MMA1_CYC, start, end, [self.v_empty[sv], self.p_free[b2], self.o_full[ob]])
            yield ("delay", 60)
            b2 += 1
            if b2 == self.NB:
                b2, use2 = 0, use2 ^ 1
            sv += 1
            if sv == self.SV:
                sv, phv = 0, phv ^ 1

    def epi_warp(self, g, q):
        speed = 1.0 + 0.05 * q + 0.02 * self.rng.random()
        for u in range(g, self.T, 2):
            b, sv = u % self.NB, u % self.SV
            conds = [(self.v_full[sv], (u // self.SV) & 1), (self.s_full[b], (u // self.NB) & 1)]
            yield ("wait", conds[::-1] if self.epi_guard_first else conds)
            if self.v_stage[sv] != (u, 0):
                self.fail(f"warp {g}.{q} reads the norms of tile {u} from V stage {sv} holding {self.v_stage[sv]}")
            cur = self.sp[b]
            if not cur or cur[1] != u or (cur[0] == "S" and not cur[2]):
                self.fail(f"warp {g}.{q} reads S of tile {u}: buffer {b} holds {cur}")
            yield ("delay", int(900 * speed))
            cur = self.sp[b]
            done = cur[2] if (cur[0] == "P" and isinstance(cur[2], dict)) else {}
            done[q] = True
            self.sp[b] = ("P", u, "complete" if len(done) == 4 else done)
            self.arrive(self.p_full[b])
            if u >= 2:  # drain the warpgroup's previous tile
                t = u - 2
                yield ("wait", [(self.o_full[g], (t >> 1) & 1)])
                if not (self.o[g] and self.o[g][1] == t and self.o[g][2]):
                    self.fail(f"warp {g}.{q} drains O[{g}] for tile {t}, holds {self.o[g]}")
                yield ("delay", int(250 * speed))
                self.o[g] = self.o[g][:3] + (self.o[g][3] + 1,)
                self.arrive(self.o_free[g])
        last = ((self.T - 1 - g) // 2) * 2 + g
        if self.T > g:
            yield ("wait", [(self.o_full[g], (last >> 1) & 1)])
            self.o[g] = self.o[g][:3] + (self.o[g][3] + 1,)
            self.arrive(self.o_free[g])


class SimOwn(SimSingle):
    """Generalisation of SimSingle to NWG epilogue warpgroups owning the sub-tiles u = g (mod NWG) (three for the small-d / k
    family and the register-contraction mode) and to the register-contraction mode (kv=True: no MMA2 and no O buffers; the
    epilogue warps release S[b] once it is in registers (p_free, four arrivals) and the V stage once it is read (v_empty,
    four arrivals); the V ring has its own producer, independent of the A ring)."""

    def __init__(self, T, seed, nwg=3, kv=False, **kw):
        super().__init__(T, seed, **kw)
        self.nwg, self.kv = nwg, kv
        self.o = [None] * nwg
        self.o_full = [Bar(1) for _ in range(nwg)]
        self.o_free = [Bar(4) for _ in range(nwg)]
        if kv:
            self.p_free = [Bar(4) for _ in range(self.NB)]
            self.v_empty = [Bar(4) for _ in range(self.SV)]

    def producer(self):
        if not self.kv:
            yield from super().producer()
            return
        sa, pha = 0, 1
        for u in range(self.T):
            yield ("wait", [(self.a_empty[sa], pha)])
            self.a_stage[sa] = ("loading", u)

            def landed(s=sa, t=u):
                self.a_stage[s] = t
                self.arrive(self.a_full[s])
            self.at(self.now + self.load_latency(), landed)
            sa += 1
            if sa == self.SA:
                sa, pha = 0, pha ^ 1
            yield ("delay", 10)

    def v_producer(self):
        sv, phv = 0, 1
        for u in range(self.T):
            yield ("wait", [(self.v_empty[sv], phv)])
            self.v_stage[sv] = ("loading", u)

            def vlanded(s=sv, t=u):
                self.v_stage[s] = (t, 0)
                self.arrive(self.v_full[s])
            self.at(self.now + self.load_latency(), vlanded)
            sv += 1
            if sv == self.SV:
                sv, phv = 0, phv ^ 1
            yield ("delay", 10)

    def mma1(self, par):
        if not self.kv:
            yield from super().mma1(par)
            return
        # S[b] is released by the epilogue (p_free) as soon as it is in registers: the buffer's previous content is "read"
        b1, sa = par % self.NB, par % self.SA
        use1, pha = (par // self.NB) & 1, (par // self.SA) & 1
        for t1 in range(par, self.T, 2):
            conds = [(self.p_free[b1], use1 ^ 1), (self.a_full[sa], pha)]
            if self.exact_guard:
                conds.insert(0, (self.a_empty[sa], pha ^ 1))
            yield ("wait", conds)

            def start(s=sa, b=b1, t=t1):
                if self.a_stage[s] != t:
                    self.fail(f"MMA1({t}) reads A stage {s} holding {self.a_stage[s]}")
                cur = self.sp[b]
                if cur is not None and not (cur[0] == "R" and cur[1] == t - self.NB and cur[2] == 4):
                    self.fail(f"MMA1({t}) overwrites S buffer {b} holding {cur}")
                self.sp[b] = ("S", t, False)

            def end(s=sa, b=b1, t=t1):
                if self.a_stage[s] != t:
                    self.fail(f"A stage {s} changed under MMA1({t}): {self.a_stage[s]}")
                self.sp[b] = ("S", t, True)
            self.issue(("mma1", par), MMA1_CYC // 2, start, end, [self.s_full[b1], self.a_empty[sa]])
            yield ("delay", 60)
            b1 += 2
            if b1 >= self.NB:
                b1, use1 = b1 - self.NB, use1 ^ 1
            sa += 2
            if sa >= self.SA:
                sa, pha = sa - self.SA, pha ^ 1

    def mma2(self):
        if self.kv:
            return
            yield
        b2 = sv = 0
        use2 = phv = 0
        for u in range(self.T):
            ob, opar = u % self.nwg, (u // self.nwg) & 1
            yield ("wait", [(self.p_full[b2], use2), (self.v_full[sv], phv), (self.o_free[ob], opar ^ 1)])

            def start(b=b2, s=sv, t=u, o=ob):
                cur = self.sp[b]
                if not (cur and cur[0] == "P" and cur[1] == t and cur[2] == "complete"):
                    self.fail(f"MMA2({t}) reads P' buffer {b} holding {cur}")
                if self.v_stage[s] != (t, 0):
                    self.fail(f"MMA2({t}) reads V stage {s} holding {self.v_stage[s]}")
                if self.o[o] is not None and self.o[o][3] != 4:
                    self.fail(f"MMA2({t}) overwrites O[{o}] = {self.o[o]} before it was drained")
                self.o[o] = ("O", t, False, 0)

            def end(b=b2, t=u, o=ob):
                self.o[o] = ("O", t, True, 0)
                self.sp[b] = ("P", t, "consumed")
            self.issue("mma2", MMA1_CYC // 2, start, end, [self.v_empty[sv], self.p_free[b2], self.o_full[ob]])
            yield ("delay", 60)
            b2 += 1
            if b2 == self.NB:
                b2, use2 = 0, use2 ^ 1
            sv += 1
            if sv == self.SV:
                sv, phv = 0, phv ^ 1

    def epi_warp(self, g, q):
        speed = 1.0 + 0.05 * q + 0.02 * self.rng.random()
        n = self.nwg
        for u in range(g, self.T, n):
            b, sv = u % self.NB, u % self.SV
            conds = [(self.v_full[sv], (u // self.SV) & 1), (self.s_full[b], (u // self.NB) & 1)]
            conds = conds[::-1] if self.epi_guard_first else conds
            if self.epi_exact:  # the buffer's previous tile has been consumed (the barrier MMA1 of this tile waits on)
                conds.insert(0, (self.p_free[b], ((u // self.NB) & 1) ^ 1))
            yield ("wait", conds)
            if self.v_stage[sv] != (u, 0):
                self.fail(f"warp {g}.{q} reads the norms of tile {u} from V stage {sv} holding {self.v_stage[sv]}")
            cur = self.sp[b]
            if not cur or cur[1] != u or (cur[0] == "S" and not cur[2]) or cur[0] == "R" and not self.kv:
                self.fail(f"warp {g}.{q} reads S of tile {u}: buffer {b} holds {cur}")
            if self.kv:
                yield ("delay", int(60 * speed))
                cur = self.sp[b]
                cnt = cur[2] + 1 if cur[0] == "R" else 1
                self.sp[b] = ("R", u, cnt)
                self.arrive(self.p_free[b])
                yield ("delay", int(500 * speed))
                if self.v_stage[sv] != (u, 0):
                    self.fail(f"V stage {sv} changed under warp {g}.{q} (tile {u}): {self.v_stage[sv]}")
                self.arrive(self.v_empty[sv])
                continue
            yield ("delay", int(600 * speed))
            cur = self.sp[b]
            done = cur[2] if (cur[0] == "P" and isinstance(cur[2], dict)) else {}
            done[q] = True
            self.sp[b] = ("P", u, "complete" if len(done) == 4 else done)
            self.arrive(self.p_full[b])
            if u >= n:
                t = u - n
                yield ("wait", [(self.o_full[g], (t // n) & 1)])
                if not (self.o[g] and self.o[g][1] == t and self.o[g][2]):
                    self.fail(f"warp {g}.{q} drains O[{g}] for tile {t}, holds {self.o[g]}")
                yield ("delay", int(150 * speed))
                self.o[g] = self.o[g][:3] + (self.o[g][3] + 1,)
                self.arrive(self.o_free[g])
        if not self.kv and self.T > g:
            last = ((self.T - 1 - g) // n) * n + g
            yield ("wait", [(self.o_full[g], (last // n) & 1)])
            self.o[g] = self.o[g][:3] + (self.o[g][3] + 1,)
            self.arrive(self.o_free[g])


def run_own(seeds, T, verbose=True, epi_exact=False, **kw):
    bad = 0
    for seed in range(seeds):
        sim = SimOwn(T, seed, **kw)
        sim.epi_exact = epi_exact
        agents = [sim.producer(), sim.mma1(0), sim.mma1(1)]
        agents.append(sim.v_producer() if sim.kv else sim.mma2())
        for a in agents:
            sim.spawn(a)
        for g in range(sim.nwg):
            for q in range(4):
                sim.spawn(sim.epi_warp(g, q))
        try:
            sim.run()
        except Violation as e:
            bad += 1
            if verbose and bad <= 3:
                print(f"  seed {seed}: {e}")
    if verbose:
        print(f"owner-drains kernel {kw} epi_exact={epi_exact}: {bad} of {seeds} runs violate an invariant (T = {T})")
    return bad


def run_single(seeds, T, verbose=True, **kw):
    bad = 0
    for seed in range(seeds):
        sim = SimSingle(T, seed, **kw)
        for agent in (sim.producer(), sim.mma1(0), sim.mma1(1), sim.mma2()):
            sim.spawn(agent)
        for g in range(2):
            for q in range(4):
                sim.spawn(sim.epi_warp(g, q))
        try:
            sim.run()
        except Violation as e:
            bad += 1
            if verbose and bad <= 4:
                print(f"  seed {seed}: {e}")
    if verbose:
        print(f"one-chunk kernel {kw}: {bad} of {seeds} runs violate an invariant (T = {T})")
    return bad


def run(schedule, seeds, T, ooo=False, verbose=True, **kw):
    bad = 0
    for seed in range(seeds):
        sim = Sim(T, schedule, seed, ooo=ooo, **kw)
        sim.spawn(sim.producer())
        sim.spawn(sim.mma1(0))
        sim.spawn(sim.mma1(1))
        sim.spawn(sim.mma2())
        for g in range(2):
            for q in range(4):
                sim.spawn(sim.epi_warp(g, q))
        try:
            sim.run()
        except Violation as e:
            bad += 1
            if verbose and bad <= 5:
                print(f"  seed {seed}: {e}")
    if verbose:
        print(f"schedule {schedule}{' (out-of-order pipe)' if ooo else ''} {kw if kw else ''}: {bad} of {seeds} runs violate an "
              f"invariant (T = {T})")
    return bad


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    ooo = "--ooo" in sys.argv
    scheds = [args[0]] if args else ["late", "early", "defer"]
    seeds = int(args[1]) if len(args) > 1 else 200
    T = int(args[2]) if len(args) > 2 else 40
    limit = next((int(a.split("=")[1]) for a in sys.argv if a.startswith("--limit=")), 0)
    if args and args[0] == "single":
        for egf in (False, True):
            for xg in (False, True):
                run_single(seeds, T, try_wait_limit=limit, epi_guard_first=egf, exact_guard=xg)
        sys.exit(0)
    for s in scheds:
        if limit:
            for gf in (False, True):
                run(s, seeds, T, ooo=ooo, try_wait_limit=limit, guard_first=gf)
        else:
            run(s, seeds, T, ooo=ooo)
