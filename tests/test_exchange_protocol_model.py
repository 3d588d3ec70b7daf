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


def simulate_rounds(nranks, ncalls, nrounds, paced, seed):
    """The same protocol with every exchange cut in `nrounds` rounds (runtime.cpp, begin_peer /
    finish_peer): sequence number of round k of call e = e * nrounds + k + 1;
      compute stream: join(e-1); for k: [paced ranks, k > 0: pace(e, k-1) = wait for every rank's round
                      k-1] pack(e,k) signal(e,k)
      comm stream   : for k: wait(e,k), which is only launched behind this rank's own signal(e,k)
      aux stream    : local(e); for k: unpack(e,k) after wait(e,k)
    Checks: no deadlock, a half is not overwritten before / while its previous reader reads it, and no
    round is unpacked before all its senders packed it."""
    rng = random.Random(seed)
    R = nrounds
    prog = {"compute": [], "comm": [], "aux": []}
    for e in range(ncalls):
        prog["compute"].append(("join", e - 1, 0))
        for k in range(R):
            if k > 0:
                prog["compute"].append(("pace", e, k - 1))
            prog["compute"] += [("pack", e, k), ("signal", e, k)]
            prog["comm"].append(("wait", e, k))
        prog["aux"].append(("local", e, 0))
        for k in range(R):
            prog["aux"] += [("unpack_begin", e, k), ("unpack_end", e, k)]
    pc = {(r, s): 0 for r in range(nranks) for s in prog}
    flags = [[0] * nranks for _ in range(nranks)]
    signalled = [0] * nranks          # own sequence numbers raised (gates this rank's wait kernels)
    packed = [[0] * nranks for _ in range(nranks)]   # packed[s][q]: sequence number of the last round packed s->q
    waited = [0] * nranks             # sequence number of the last completed wait
    unpacked = [0] * nranks
    reading = [None] * nranks
    started = [-1] * nranks
    violations = []
    seq = lambda e, k: e * R + k + 1

    def enabled(r, s):
        i = pc[(r, s)]
        if i >= len(prog[s]):
            return False
        op, e, k = prog[s][i]
        if op == "join":
            return e < 0 or (unpacked[r] >= seq(e, R - 1) and waited[r] >= seq(e, R - 1))
        if op == "pace":
            return r not in paced or all(flags[r][x] >= seq(e, k) for x in range(nranks))
        if op in ("pack", "signal", "unpack_end"):
            return True
        if op == "wait":
            return signalled[r] >= seq(e, k) and all(flags[r][x] >= seq(e, k) for x in range(nranks))
        if op == "local":
            return started[r] >= e
        if op == "unpack_begin":
            return waited[r] >= seq(e, k)
        raise AssertionError(op)

    def step(r, s):
        i = pc[(r, s)]
        op, e, k = prog[s][i]
        if op == "join":
            started[r] = e + 1
        elif op == "pack":
            for q in range(nranks):
                if e >= 2 and unpacked[q] < seq(e - 2, R - 1):
                    violations.append(("overwrite before read", r, q, e, k))
                # (rounds of the SAME call use disjoint windows of the half: only another call's reader counts)
                if reading[q] is not None and reading[q] != e and (reading[q] & 1) == (e & 1):
                    violations.append(("overwrite while reading", r, q, e, k))
                packed[r][q] = seq(e, k)
        elif op == "signal":
            signalled[r] = seq(e, k)
            for q in range(nranks):
                flags[q][r] = seq(e, k)
        elif op == "wait":
            waited[r] = seq(e, k)
        elif op == "unpack_begin":
            for x in range(nranks):
                if packed[x][r] < seq(e, k):
                    violations.append(("read before written", x, r, e, k))
            reading[r] = e
        elif op == "unpack_end":
            reading[r] = None
            unpacked[r] = seq(e, k)
        pc[(r, s)] = i + 1

    while True:
        ready = [(r, s) for r in range(nranks) for s in prog if enabled(r, s)]
        if not ready:
            break
        r, s = rng.choice(ready) if rng.random() < 0.7 else min(ready, key=lambda x: (x[0] + seed) % nranks)
        step(r, s)
    done = all(pc[(r, s)] == len(prog[s]) for r in range(nranks) for s in prog)
    return done, violations


@pytest.mark.parametrize("nranks,paced", [(2, ()), (4, (0, 3)), (8, (0, 7)), (3, (0, 1, 2))])
def test_rounds_gated_waits_and_paced_senders(nranks, paced):
    """Rounds, wait kernels gated on the rank's own signal, and senders that pace themselves on the
    previous round of every rank (even all of them): no deadlock, no half overwritten early."""
    for seed in range(150):
        done, violations = simulate_rounds(nranks, 5, 3, set(paced), seed)
        assert done, "deadlock in the model"
        assert not violations, violations[:3]
