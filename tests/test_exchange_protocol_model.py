"""Model check of the peer-memory exchange protocol (superbblas_b200/csrc/runtime.cpp, execute_copy;
DESIGN.md §5) without a GPU: a randomised interleaving of the stream operations of R ranks over
several consecutive exchanges, checking the property the alternation of the arena halves relies on.

Per rank and call e (half h = e & 1), in stream order:
  compute stream: [join of call e-1]  pack(e): stores into every peer's arena half h, then signal(e)
  comm stream   : wait(e): completes when every rank's flag has reached e
  aux stream    : starts after the compute stream reached call e; local(e); after wait(e): unpack(e),
                  which reads this rank's arena half h
  end of call   : the compute stream continues after unpack(e) and wait(e)            (join of call e)
Property: when a rank packs call e into a peer's half h, that peer has finished unpack(e-2), the
previous reader of that half; and nobody unpacks a half before all its writers are done.
The second test shows that the property fails if ranks only wait for the peers they receive from
(why every rank signals every rank in every round)."""
import random

import pytest


def simulate(nranks, ncalls, seed, all_to_all=True, receivers=None):
    rng = random.Random(seed)
    # state per rank
    pc = {(r, s): 0 for r in range(nranks) for s in ("compute", "comm", "aux")}
    flags = [[0] * nranks for _ in range(nranks)]       # flags[q][r]: what rank r signalled to rank q
    packed = [[-1] * nranks for _ in range(nranks)]      # packed[s][q]: last call whose pack s->q completed
    unpacked = [-1] * nranks                             # last call whose unpack completed
    waited = [-1] * nranks
    reading = [None] * nranks                            # call whose unpack is in progress
    violations = []
    # programs: compute: for e: ("join", e-1), ("pack", e), ("signal", e); comm: ("wait", e); aux: ("local", e), ("unpack", e)
    prog = {
        "compute": [op for e in range(ncalls) for op in (("join", e - 1), ("pack", e), ("signal", e))],
        "comm": [("wait", e) for e in range(ncalls)],
        "aux": [op for e in range(ncalls) for op in (("local", e), ("unpack_begin", e), ("unpack_end", e))],
    }
    started = [-1] * nranks  # call whose start the compute stream has reached (ev_a)

    def senders_to(q, e):
        return range(nranks) if receivers is None else [s for s in range(nranks) if q in receivers(s, e)]

    def enabled(r, s):
        i = pc[(r, s)]
        if i >= len(prog[s]):
            return False
        op, e = prog[s][i]
        if op == "join":
            return e < 0 or (unpacked[r] >= e and waited[r] >= e)
        if op in ("pack", "signal"):
            return True
        if op == "wait":
            need = range(nranks) if all_to_all else senders_to(r, e)
            return started[r] >= e and all(flags[r][x] >= e + 1 for x in need)
        if op == "local":
            return started[r] >= e
        if op == "unpack_begin":
            return waited[r] >= e
        if op == "unpack_end":
            return True
        raise AssertionError(op)

    def step(r, s):
        i = pc[(r, s)]
        op, e = prog[s][i]
        h = e & 1
        if op == "join":
            started[r] = e + 1
        elif op == "pack":
            for q in (range(nranks) if receivers is None else receivers(r, e)):
                # the previous reader of half h at q is its unpack of call e-2
                if e >= 2 and unpacked[q] < e - 2:
                    violations.append(("overwrite before read", r, q, e))
                if reading[q] is not None and (reading[q] & 1) == h:
                    violations.append(("overwrite while reading", r, q, e))
                packed[r][q] = e
        elif op == "signal":
            for q in range(nranks):
                flags[q][r] = e + 1
        elif op == "wait":
            waited[r] = e
        elif op == "unpack_begin":
            for x in senders_to(r, e):
                if packed[x][r] < e:
                    violations.append(("read before written", x, r, e))
            reading[r] = e
        elif op == "unpack_end":
            reading[r] = None
            unpacked[r] = e
        pc[(r, s)] = i + 1

    for r in range(nranks):
        started[r] = -1
    while True:
        ready = [(r, s) for r in range(nranks) for s in ("compute", "comm", "aux") if enabled(r, s)]
        if not ready:
            break
        # bias the schedule so that some ranks run far ahead of others
        r, s = rng.choice(ready) if rng.random() < 0.7 else min(ready, key=lambda x: (x[0] + seed) % nranks)
        step(r, s)
    done = all(pc[(r, s)] == len(prog[s]) for r in range(nranks) for s in prog)
    return done, violations


@pytest.mark.parametrize("nranks", [2, 3, 8])
def test_halves_are_never_overwritten_before_they_are_read(nranks):
    for seed in range(300):
        done, violations = simulate(nranks, 6, seed)
        assert done, "deadlock in the model"
        assert not violations, violations[:3]


def test_waiting_only_for_actual_senders_is_not_enough():
    """Sparse exchange (every rank sends to its right neighbour only): if a rank waited only for the
    ranks it receives from, a fast sender could overwrite a half its receiver has not read yet."""
    ring = lambda s, e: [(s + 1) % 3]
    bad = 0
    for seed in range(300):
        done, violations = simulate(3, 6, seed, all_to_all=False, receivers=ring)
        assert done
        bad += bool(violations)
    assert bad > 0
    for seed in range(300):
        done, violations = simulate(3, 6, seed, all_to_all=True, receivers=ring)
        assert done and not violations
