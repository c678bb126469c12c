"""Barrier protocol of the two-chunk tensor-core kernel (sliced epilogue), checked on the discrete-event model of
scripts/tc_protocol_model.py: mbarrier phase / parity semantics (with their two-phase ambiguity), ring depths and arrival
counts as in rlaopt_b200/csrc/kmm_tc.cu, random bulk-copy latencies with spikes.  CPU only."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
import tc_protocol_model as M  # noqa: E402


def _run(cls, schedule, seed, T=24, ooo=False):
    sim = cls(T, schedule, seed, ooo=ooo, exact_guard=True)
    for agent in (sim.producer(), sim.mma1(0), sim.mma1(1), sim.mma2()):
        sim.spawn(agent)
    for g in range(2):
        for q in range(4):
            sim.spawn(sim.epi_warp(g, q))
    sim.run()


@pytest.mark.parametrize("schedule", ["late", "early", "defer"])
@pytest.mark.parametrize("ooo", [False, True])
def test_schedules_keep_every_data_invariant(schedule, ooo):
    """No ordering the barriers allow lets MMA1 read a foreign or half-landed image, the pointwise stage an incomplete S,
    MMA2 an incomplete P' or a foreign V record, or a drain an O buffer that is being overwritten -- for the shipped
    schedule and for the two experimental ones (so the mis-computation of the early schedule on the hardware,
    profiles/r02_tc_dual_sliced.md, is not a protocol error the model can see)."""
    for seed in range(12):
        _run(M.Sim, schedule, seed, ooo=ooo)


def test_model_sees_a_missing_guard():
    """The A ring is three stages deep and shared by the two MMA1 issue warps, so an a_full parity wait alone is ambiguous
    (the other warp's tile may not have landed yet): it is the simultaneous p_free wait that makes it safe.  Without it
    the model finds the stale-image read at once."""
    class NoGuard(M.Sim):
        def mma1(self, par):
            for req in super().mma1(par):
                yield ("wait", req[1][:1]) if req[0] == "wait" else req

    hits = 0
    for seed in range(12):
        try:
            _run(NoGuard, "late", seed)
        except M.Violation as e:
            assert "MMA1" in str(e)
            hits += 1
    assert hits > 0


def test_model_sees_a_wrong_arrival_count():
    class HalfCount(M.Sim):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.o_free = [M.Bar(4) for _ in range(2)]

    with pytest.raises(M.Violation):
        _run(HalfCount, "late", 0)


def _run_kw(schedule, seed, **kw):
    kw.setdefault("exact_guard", False)  # the two-chunk plan has NB = SA = 3: p_free alone is a sufficient guard there
    sim = M.Sim(24, schedule, seed, **kw)
    for agent in (sim.producer(), sim.mma1(0), sim.mma1(1), sim.mma2()):
        sim.spawn(agent)
    for g in range(2):
        for q in range(4):
            sim.spawn(sim.epi_warp(g, q))
    sim.run()


@pytest.mark.parametrize("schedule", ["late", "early", "defer"])
def test_guard_first_poll_order_is_safe_under_suspending_try_wait(schedule):
    """mbarrier.try_wait may suspend the thread up to a system time limit, so the barriers of one multi-barrier poll are
    sampled at different times.  The MMA1 issue warps poll p_free (the guard) BEFORE a_full (whose parity is ambiguous
    while the other issue warp's image is in flight): safe for any suspension time."""
    for limit in (500, 8000, 30000):
        for seed in range(8):
            _run_kw(schedule, seed, try_wait_limit=limit, guard_first=True)


def test_a_full_first_poll_order_is_not():
    """The order the kernel had before: an a_full sample taken while the other warp's image was still in flight reads
    "complete", and if the p_free wait that follows returns within one suspension MMA1 consumes a stage whose bulk copy
    has only just been issued."""
    hits = 0
    for seed in range(30):
        try:
            _run_kw("late", seed, try_wait_limit=20000, guard_first=False)
        except M.Violation as e:
            assert "MMA1" in str(e) and "A stage" in str(e)
            hits += 1
    assert hits > 0


# ---- the one-chunk kernels (two epilogue warpgroups that own alternate sub-tiles), every ring plan tc_plan produces -------
PLANS_NWG2 = [(2, 2, 4), (3, 3, 4), (4, 3, 5), (4, 4, 6)]  # (NB, SA, SV); {2, 2, 3} is bumped to {2, 2, 4} by the plan


@pytest.mark.parametrize("nb,sa,sv", PLANS_NWG2)
def test_one_chunk_plans_keep_every_data_invariant(nb, sa, sv):
    assert M.run_single(40, 40, verbose=False, nb=nb, sa=sa, sv=sv, exact_guard=True) == 0


def test_c2_plan_needs_the_a_empty_guard():
    """NB = 4, SA = 3 (d = 128, k = 64): the two MMA1 issue warps share the three A stages, the a_full parity wait is
    ambiguous while the other warp's tile t - 3 is in flight, and p_free(t - 4) does not cover it.  With the third
    condition -- a_empty of the stage's previous tile, polled first -- the hole is closed (the kernel's wait)."""
    assert M.run_single(40, 40, verbose=False, nb=4, sa=3, sv=5, exact_guard=False) > 0
    assert M.run_single(40, 40, verbose=False, nb=4, sa=3, sv=5, exact_guard=True) == 0
    assert M.run_single(40, 40, verbose=False, nb=4, sa=3, sv=5, exact_guard=True, try_wait_limit=20000) == 0


def test_odd_v_ring_of_the_k128_plan_is_why_it_got_a_fourth_stage():
    """{NB 2, SA 2, SV 3}: a warpgroup's v_full wait is ambiguous (odd ring shared with the other warpgroup) and sampled
    before its s_full guard; with suspending try_waits the pointwise stage can read the norms of a record in flight.
    {2, 2, 4} has no ambiguous wait."""
    assert M.run_single(60, 40, verbose=False, nb=2, sa=2, sv=3, exact_guard=True, try_wait_limit=20000) > 0
    assert M.run_single(60, 40, verbose=False, nb=2, sa=2, sv=4, exact_guard=True, try_wait_limit=20000) == 0


# ---- three epilogue warpgroups (small d / k family, register-contraction mode): the plans tc_plan produces --------------
@pytest.mark.parametrize("kv,nb,sa,sv", [(False, 4, 4, 6), (False, 5, 4, 6), (True, 5, 4, 6)])
@pytest.mark.parametrize("limit", [0, 20000])
def test_three_warpgroup_plans_keep_every_data_invariant(kv, nb, sa, sv, limit):
    assert M.run_own(30, 48, verbose=False, nwg=3, kv=kv, nb=nb, sa=sa, sv=sv, exact_guard=True, try_wait_limit=limit) == 0


def test_register_contraction_with_three_k_blocks_is_why_it_stops_at_d_128():
    """{NB 4, SA 3, SV 5} with three warpgroups (what 128 < d <= 192, k <= 4 would get): both epilogue waits are ambiguous
    and unguarded -- stale S and stale norms.  tc_plan sends those shapes to the two-warpgroup MMA2 kernels."""
    assert M.run_own(30, 48, verbose=False, nwg=3, kv=True, nb=4, sa=3, sv=5, exact_guard=True) > 0
    assert M.run_own(30, 48, verbose=False, nwg=2, kv=False, nb=4, sa=3, sv=5, exact_guard=True) == 0


@pytest.mark.parametrize("nb,sa,sv", [(2, 2, 4), (3, 3, 4)])  # 64 < d <= 128 (default) and d <= 64 (RLAOPT_B200_TC_DUAL=2)
@pytest.mark.parametrize("limit", [0, 20000])
def test_drains_last_two_chunk_kernels_keep_every_data_invariant(nb, sa, sv, limit):
    assert M.run("last", 30, 40, verbose=False, nb=nb, sa=sa, sv=sv, exact_guard=True, try_wait_limit=limit) == 0
