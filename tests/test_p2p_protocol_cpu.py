"""Model check of the peer-memory redistribution protocol (elemental_b200/csrc/host/redist.cpp,
kernels/p2p.cu) under random interleavings of the ranks.

Per epoch e of a channel every rank runs, in stream order:
    PUSH(e)  write its pieces into region (e & 1, me) of each destination's window
    XCHG(e)  ready[dest][me] := e for each destination; wait ready[me][src] >= e for each source
    UNPK(e)  read region (e & 1, src) of its own window for each source
    ACK(e)   ack[w][me] := e for every peer w; wait ack[me][w] >= e - 1 for every peer w
The model executes one step of a randomly chosen runnable rank at a time (a waiting step is runnable
only when its condition holds) and checks
  * safety: UNPK(e) finds exactly the epoch-e data of every source (never a stale e-2 piece, never one
    already overwritten by e+2), and PUSH never overwrites a piece its destination has not consumed;
  * liveness: some rank is always runnable until all have finished (no deadlock), whatever the
    communication pattern of each epoch, including ranks that neither send nor receive.
A weakened protocol (ACK waits dropped) must be caught, which shows the check has teeth.
"""
import random

import pytest


def simulate(p, epochs, seed, ack_wait=True, max_steps=10**6):
    rng = random.Random(seed)
    # dests[e][r] = set of ranks r pushes to in epoch e (any pattern, empty sets allowed)
    dests = [None] + [[set(d for d in range(p) if d != r and rng.random() < 0.5) for r in range(p)]
                      for _ in range(epochs)]
    srcs = [None] + [[{r for r in range(p) if d in dests[e][r]} for d in range(p)] for e in range(1, epochs + 1)]
    ready = [[0] * p for _ in range(p)]      # ready[owner][src]
    ack = [[0] * p for _ in range(p)]        # ack[owner][src]
    window = [[[0] * p for _ in range(2)] for _ in range(p)]     # window[owner][half][src] = epoch of the data
    consumed = [[[0] * p for _ in range(2)] for _ in range(p)]   # last epoch the owner has unpacked from there
    pc = [(1, 0)] * p                        # (epoch, step) with step 0 PUSH, 1 XCHG-signal, 2 XCHG-wait, 3 UNPK, 4 ACK-signal, 5 ACK-wait
    done = [False] * p
    steps = 0
    while not all(done):
        runnable = []
        for r in range(p):
            if done[r]:
                continue
            e, st = pc[r]
            if st == 2 and not all(ready[r][s] >= e for s in srcs[e][r]):
                continue
            if st == 5 and ack_wait and not all(ack[r][w] >= e - 1 for w in range(p) if w != r):
                continue
            runnable.append(r)
        if not runnable:
            return "deadlock"
        r = rng.choice(runnable)
        e, st = pc[r]
        if st == 0:
            for d in dests[e][r]:
                prev = window[d][e & 1][r]
                if prev and consumed[d][e & 1][r] < prev:
                    return f"rank {r} overwrote epoch {prev} in rank {d}'s window before it was consumed"
                window[d][e & 1][r] = e
        elif st == 1:
            for d in dests[e][r]:
                ready[d][r] = e
        elif st == 3:
            for s in srcs[e][r]:
                if window[r][e & 1][s] != e:
                    return f"rank {r} unpacked epoch {window[r][e & 1][s]} instead of {e} from rank {s}"
                consumed[r][e & 1][s] = e
        elif st == 4:
            for w in range(p):
                if w != r:
                    ack[w][r] = e
        st += 1
        if st == 6:
            e, st = e + 1, 0
            if e > epochs:
                done[r] = True
        pc[r] = (e, st)
        steps += 1
        if steps > max_steps:
            return "step limit"
    return "ok"


@pytest.mark.parametrize("p", [2, 3, 4, 8])
def test_protocol_is_safe_and_live_under_random_interleavings(p):
    for seed in range(60):
        assert simulate(p, epochs=12, seed=seed) == "ok", (p, seed)


def test_dropping_the_ack_wait_is_caught():
    outcomes = {simulate(4, epochs=12, seed=s, ack_wait=False) for s in range(60)}
    assert any(o != "ok" for o in outcomes), "the model does not detect a window half reused too early"
