"""CPU: randomized model check of the operand-ring protocol of csrc/sparse_conv_g4.cu (producers, MMA warps, parity waits on mbarriers,
asynchronous MMA completion) -- for the production hand-off (ONE turn counter PER RING SLOT: a stage waits until the previous use of
its slot has been seen full, so the parity wait is unambiguous while stages on different slots do not wait for each other) and for
the earlier global hand-offs (one turn counter over all stages, passed on after issuing / as soon as the slot was seen full).

What is modelled (one CTA, one pass):
  * stage s lives in ring slot s % NA with phase parity (s / NA) & 1; full[slot] counts the producers' arrivals, empty[slot] the
    completion of the stage's MMAs (tcgen05.commit);
  * mbarrier.try_wait.parity(p) succeeds when the barrier's current phase parity differs from p (the phase with parity p completed);
    a fresh barrier therefore lets a wait on parity 1 pass, which is how the producers' first use of a slot does not block;
  * a producer owns one slot and fills it for every NA-th stage after waiting for empty[slot]; an MMA warp walks ALL stages in order
    and acts on those whose sub-tile it owns, after waiting for its turn and for full[slot]; MMAs execute later, FIFO per issuing
    warp but in any order between warps, read the slot when they execute and arrive on empty[slot] when they complete.
Checked under random interleavings: no deadlock, every stage's MMAs read the slot contents of exactly that stage, and every stage is
executed once.  (The weight-slab ring and the TMEM accumulators are not modelled: the hand-off change does not touch them.)"""
import random

import pytest


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.parity = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier expects in one phase"
        if self.pending == 0:
            self.pending, self.parity = self.count, self.parity ^ 1

    def try_wait(self, parity):
        return self.parity != parity


def simulate(n_stages, n_slots, n_sub, early_turn, seed, halves=2, n_mma_warps=4, producers_wait_for_empty=True, turn="global"):
    rng = random.Random(seed)
    full = [MBar(halves) for _ in range(n_slots)]
    empty = [MBar(1) for _ in range(n_slots)]
    slot_tag = [None] * n_slots          # stage whose rows the slot currently holds (None = garbage)
    slot_fill = [0] * n_slots            # producers that have finished copying the current stage
    state = {"turn": 0}
    slot_turn = [0] * n_slots            # turn == "slot": uses of the slot that have been seen full
    executed = []
    queues = [[] for _ in range(n_mma_warps)]          # issued, not yet executed MMAs per issuing warp

    def producer(slot, half):
        for s in range(slot, n_stages, n_slots):
            phase = (s // n_slots) & 1
            while producers_wait_for_empty and not empty[slot].try_wait(phase ^ 1):
                yield
            for _ in range(rng.randint(0, 3)):          # the copies take a while
                yield
            slot_fill[slot] += 1
            if slot_fill[slot] == halves:
                slot_tag[slot], slot_fill[slot] = s, 0
            full[slot].arrive()
            yield

    def mma_warp(w):
        for s in range(n_stages):
            slot, phase, j = s % n_slots, (s // n_slots) & 1, s % n_sub
            if (j % n_mma_warps) != w:
                continue
            for _ in range(rng.randint(0, 2)):          # per-stage bookkeeping (iterator, descriptors)
                yield
            if turn == "slot":
                while slot_turn[slot] != s // n_slots:
                    yield
            elif turn == "global":
                while state["turn"] != s:
                    yield
            while not full[slot].try_wait(phase):
                if turn == "none":
                    yield
                    continue
                yield
            if turn == "slot":
                slot_turn[slot] = s // n_slots + 1
                yield
            if early_turn:
                state["turn"] = s + 1
                yield
            queues[w].append((s, slot))                  # the stage's MMAs are issued
            for _ in range(rng.randint(0, 2)):
                yield
            if not early_turn:
                state["turn"] = s + 1
            yield

    def tensor_pipe():
        while len(executed) < n_stages:
            ready = [q for q in queues if q]
            if ready:
                s, slot = rng.choice(ready).pop(0)
                assert slot_tag[slot] == s, f"stage {s} read slot {slot} holding stage {slot_tag[slot]}"
                executed.append(s)
                empty[slot].arrive()                     # tcgen05.commit -> empty[slot]
            yield

    procs = [producer(k, h) for k in range(n_slots) for h in range(halves)] + [mma_warp(w) for w in range(n_mma_warps)] + [tensor_pipe()]
    alive = list(procs)
    for _ in range(400 * n_stages + 10000):
        if not alive:
            break
        p = rng.choice(alive)
        try:
            next(p)
        except StopIteration:
            alive.remove(p)
    assert not alive, f"deadlock: {len(executed)} of {n_stages} stages executed"
    assert sorted(executed) == list(range(n_stages))
    return executed


@pytest.mark.parametrize("early_turn", [False, True])
@pytest.mark.parametrize("n_slots,n_sub", [(5, 3), (5, 1), (4, 2), (8, 8), (5, 4), (8, 5)])
def test_ring_protocol_under_random_interleavings(early_turn, n_slots, n_sub):
    for seed in range(40):
        simulate(n_stages=37 + seed % 11, n_slots=n_slots, n_sub=n_sub, early_turn=early_turn, seed=seed)


def test_the_model_detects_a_broken_protocol():
    """Sanity of the checker itself: producers that do not wait for empty[slot] overwrite rows that are still to be read."""
    caught = 0
    for seed in range(20):
        try:
            simulate(n_stages=40, n_slots=5, n_sub=3, early_turn=False, seed=seed, producers_wait_for_empty=False)
        except AssertionError:
            caught += 1
    assert caught >= 15


@pytest.mark.parametrize("n_slots,n_sub", [(5, 3), (5, 1), (4, 2), (8, 8), (5, 4), (8, 5), (6, 7)])
def test_per_slot_turn_counters_under_random_interleavings(n_slots, n_sub):
    """The production hand-off: a stage only waits for the previous use of ITS slot."""
    for seed in range(60):
        simulate(n_stages=37 + seed % 23, n_slots=n_slots, n_sub=n_sub, early_turn=False, seed=seed, turn="slot")


def test_without_any_turn_the_parity_wait_is_ambiguous():
    """Why a hand-off is needed at all: a warp that reaches a later use of a slot while an earlier use is still pending sees the
    barrier's stale parity and issues on rows that are not there yet."""
    caught = 0
    for seed in range(30):
        try:
            simulate(n_stages=60, n_slots=5, n_sub=3, early_turn=False, seed=seed, turn="none", n_mma_warps=4)
        except AssertionError:
            caught += 1
    assert caught >= 20
