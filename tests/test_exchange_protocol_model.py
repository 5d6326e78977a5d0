"""Exhaustive model check of the Lamport-style gradient exchange of dwopt_kernel (minppo_b200/csrc/dwopt.cuh).

The CUDA protocol has no flags, fences or acknowledgements: staging words hold a sentinel until a peer's store lands,
the consumer polls, sums, and puts the sentinel back; slots alternate with the PARITY of the exchange number.  What
makes two slots enough is an ordering argument (a rank can only start exchange n+1 after it completed exchange n, and
its clears are performed before its next launch).  This test checks that argument by brute force: a small abstract
machine -- W ranks, remote stores delivered after an ARBITRARY delay and in ANY order -- is explored over every
interleaving, for both variants (one-shot: everybody pushes to everybody; two-phase: unit u is reduced by rank u, which
pushes the result back).  Violations looked for: a store landing on a word that still holds unread data, a consumer
reading a value of the wrong exchange, a deadlock.  The same machine with ONE slot per source (no parity) must fail --
otherwise the checker would prove nothing."""
import itertools

import pytest

SENT = None


def _explore(W, exchanges, nslots, two_phase):
    """DFS over all interleavings.  Returns (ok, reason).  One unit per rank (unit u owned by rank u in two-phase).

    Memory of rank r: stage[(slot, src, unit)] and result[(slot, unit)].  Value written = (src_or_owner, n, unit).
    A rank's exchange n consists of sub-tasks; it may start exchange n+1 only when all are done (kernel boundary)."""

    def initial():
        ranks = tuple((0, start_tasks(r)) for r in range(W))
        return (ranks, frozenset(), tuple(frozenset() for _ in range(W)))      # (ranks, in-flight, memory per rank)

    def start_tasks(r):
        if not two_phase:
            return frozenset({("push",), ("reduce",)})
        t = {("own_reduce",)}
        for o in range(W):
            if o != r:
                t.add(("contrib", o))
                t.add(("wait_result", o))
        return frozenset(t)

    def mem_get(mem, key):
        for k, v in mem:
            if k == key:
                return v
        return SENT

    def mem_set(mem, key, val):
        d = dict(mem)
        if val is SENT:
            d.pop(key, None)
        else:
            d[key] = val
        return frozenset(d.items())

    seen = set()
    stack = [initial()]
    while stack:
        state = stack.pop()
        if state in seen:
            continue
        seen.add(state)
        ranks, flight, mems = state
        if all(n == exchanges for n, _ in ranks):
            if flight:
                return False, "stores still in flight after the last exchange"
            continue
        succ = []
        # (1) deliver any in-flight store
        for msg in flight:
            dst, key, val = msg
            if mem_get(mems[dst], key) is not SENT:
                return False, f"store {val} landed on unread data {mem_get(mems[dst], key)} at rank {dst} {key}"
            m2 = list(mems)
            m2[dst] = mem_set(mems[dst], key, val)
            succ.append((ranks, flight - {msg}, tuple(m2)))
        # (2) any rank performs one enabled sub-task
        for r, (n, tasks) in enumerate(ranks):
            if n == exchanges:
                continue
            slot = n % nslots
            for task in tasks:
                new_flight, mem_r, done = flight, mems[r], False
                if task == ("push",):                                        # one-shot: my sums, to every peer (units are
                    add = {(q, ("stage", slot, r, 0), (r, n, 0)) for q in range(W) if q != r}   # independent: one suffices)
                    new_flight, done = flight | add, True
                elif task == ("reduce",):                                    # one-shot: all peers' words of the unit
                    keys = [("stage", slot, q, 0) for q in range(W) if q != r]
                    vals = [mem_get(mem_r, k) for k in keys]
                    if all(v is not SENT for v in vals):
                        for k, v in zip(keys, vals):
                            if v != (k[2], n, k[3]):
                                return False, f"rank {r} exchange {n} read {v} from {k}"
                            mem_r = mem_set(mem_r, k, SENT)
                        done = True
                elif task[0] == "contrib":                                   # two-phase: my contribution to unit o -> owner o
                    o = task[1]
                    new_flight, done = flight | {(o, ("stage", slot, r, o), (r, n, o))}, True
                elif task == ("own_reduce",):                                # two-phase: reduce my unit, push the result
                    keys = [("stage", slot, q, r) for q in range(W) if q != r]
                    vals = [mem_get(mem_r, k) for k in keys]
                    if all(v is not SENT for v in vals):
                        for k, v in zip(keys, vals):
                            if v != (k[2], n, r):
                                return False, f"owner {r} exchange {n} read {v} from {k}"
                            mem_r = mem_set(mem_r, k, SENT)
                        new_flight = flight | {(q, ("result", slot, r), ("res", n, r)) for q in range(W) if q != r}
                        done = True
                elif task[0] == "wait_result":
                    o = task[1]
                    k = ("result", slot, o)
                    v = mem_get(mem_r, k)
                    if v is not SENT:
                        if v != ("res", n, o):
                            return False, f"rank {r} exchange {n} read result {v} of unit {o}"
                        mem_r, done = mem_set(mem_r, k, SENT), True
                if not done:
                    continue
                left = tasks - {task}
                # the sub-tasks of one exchange are threads of ONE launch: the next launch starts when all are done
                nr = (n, left) if left else ((n + 1, start_tasks(r)) if n + 1 < exchanges else (exchanges, frozenset()))
                r2 = list(ranks)
                r2[r] = nr
                m2 = list(mems)
                m2[r] = mem_r
                succ.append((tuple(r2), new_flight, tuple(m2)))
        if not succ:
            return False, "deadlock"
        stack.extend(succ)
    return True, f"{len(seen)} states"


@pytest.mark.parametrize("two_phase", [False, True])
@pytest.mark.parametrize("W", [2, 3])
def test_two_parity_slots_are_enough(W, two_phase):
    ok, why = _explore(W, exchanges=4 if (W == 2 or not two_phase) else 3, nslots=2, two_phase=two_phase)
    assert ok, why


def test_single_slot_is_not_enough_for_one_shot():
    """Without the parity double-buffer a fast rank's next push lands before the slow rank consumed the last one --
    the checker does find real violations."""
    ok, why = _explore(2, exchanges=3, nslots=1, two_phase=False)
    assert not ok and "unread" in why, why


def test_two_phase_has_a_slot_of_margin():
    """In the two-phase variant every store is causally behind the consumption of its predecessor (a contribution
    n+1 needs the result n, a result n+1 needs the contribution n+1), so it would be safe even with ONE slot; the kernel
    still alternates two."""
    ok, why = _explore(2, exchanges=3, nslots=1, two_phase=True)
    assert ok, why
