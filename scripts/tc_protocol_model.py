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
    def __init__(self, T, schedule, seed, ooo=False, spike=0.02, try_wait_limit=0, guard_first=True):
        # try_wait_limit = 0: the barriers of one poll are sampled at the same instant (idealised); > 0: sampled one after
        # the other, each mbarrier.try_wait suspending the thread until its phase completes or the limit (cycles) expires
        # -- a sample taken early in the poll can be stale when the last one returns.  guard_first: order of the two
        # barriers in the MMA1 issue warps' wait (True = p_free then a_full, the kernel since the fix; False = before).
        self.T, self.schedule, self.ooo = T, schedule, ooo
        self.L, self.guard_first = try_wait_limit, guard_first
        self.single = []
        self.rng = random.Random(seed)
        self.spike = spike
        self.now = 0
        self.events = []  # (time, seq, fn)
        self.seq = 0
        self.waiters = []  # (agent generator, condition)
        b = lambda c: Bar(c)
        self.a_full = [b(1) for _ in range(SA)]
        self.a_empty = [b(1) for _ in range(SA)]
        self.v_full = [b(1) for _ in range(SV)]
        self.v_empty = [b(1) for _ in range(SV)]
        self.s_full = [b(1) for _ in range(NB)]
        self.p_full = [b(4) for _ in range(NB)]
        self.p_free = [b(1) for _ in range(NB)]
        self.o_full = [b(1) for _ in range(2)]
        self.o_free = [b(8) for _ in range(2)]
        # contents
        self.a_stage = [None] * SA        # tile id, or ("loading", tile)
        self.v_stage = [None] * SV        # (tile, chunk) or ("loading", ...)
        self.sp = [None] * NB             # ("S", t, done) / ("P", t, quarters_done_by_warp dict)
        self.sp_busy = [None] * NB        # tensor op currently writing / reading: (kind, t, end)
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
                if sv == SV:
                    sv, phv = 0, phv ^ 1
                yield ("delay", 10)
            sa += 1
            if sa == SA:
                sa, pha = 0, pha ^ 1
            yield ("delay", 10)

    def mma1(self, par):
        b1, sa = par % NB, par % SA
        use1, pha = (par // NB) & 1, (par // SA) & 1
        for t1 in range(par, self.T, 2):
            conds = [(self.p_free[b1], use1 ^ 1), (self.a_full[sa], pha)]
            yield ("wait", conds if self.guard_first else conds[::-1])

            def start(s=sa, b=b1, t=t1):
                if self.a_stage[s] != t:
                    self.fail(f"MMA1({t}) reads A stage {s} holding {self.a_stage[s]}")
                cur = self.sp[b]
                if cur is not None and not (cur[0] == "P" and cur[1] == t - NB and cur[2] == "consumed"):
                    self.fail(f"MMA1({t}) overwrites S/P buffer {b} holding {cur}")
                self.sp[b] = ("S", t, False)

            def end(s=sa, b=b1, t=t1):
                if self.a_stage[s] != t:
                    self.fail(f"A stage {s} changed under MMA1({t}): {self.a_stage[s]}")
                self.sp[b] = ("S", t, True)
            self.issue(("mma1", par), MMA1_CYC, start, end, [self.s_full[b1], self.a_empty[sa]])
            yield ("delay", 60)
            b1 += 2
            if b1 >= NB:
                b1, use1 = b1 - NB, use1 ^ 1
            sa += 2
            if sa >= SA:
                sa, pha = sa - SA, pha ^ 1

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
                if sv == SV:
                    sv, phv = 0, phv ^ 1
            b2 += 1
            if b2 == NB:
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
            cur = self.sp[u % NB]
            if not cur or cur[1] != u or (cur[0] == "S" and not cur[2]):
                self.fail(f"warp {g}.{q} {what} of tile {u}: buffer {u % NB} holds {cur}")

        def slice0(u):
            yield ("wait", [(self.s_full[u % NB], (u // NB) & 1)])
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
            self.arrive(self.p_full[u % NB])

        def publish(u, defer):
            if defer:
                pending[0] = u
            else:
                announce(u)

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
        b = u % NB
        cur = self.sp[b]
        if cur[0] == "S":
            cur = ("P", u, {})
        done = cur[2] if isinstance(cur[2], dict) else {}
        done[(q, sl)] = True
        self.sp[b] = ("P", u, "complete" if len(done) == 16 else done)


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
    for s in scheds:
        if limit:
            for gf in (False, True):
                run(s, seeds, T, ooo=ooo, try_wait_limit=limit, guard_first=gf)
        else:
            run(s, seeds, T, ooo=ooo)
