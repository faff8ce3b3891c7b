"""CPU model of the ghost-row protocol the strip day kernel uses (nesosim_b200/csrc/day_kernels.cuh, StripLink):
per strip and per side one flag (latest slot delivered) and one mailbox double-buffered by slot parity.  A day x of a
strip is: wait for both flags >= x (x > 0), read the mailboxes of parity x & 1, compute, write the neighbours' mailboxes
of parity (x+1) & 1, raise the neighbours' flags to x+1.  The claim in DESIGN.md §6 -- no further handshake is needed
because a strip can never be more than one day ahead of a neighbour -- is checked here over random and adversarial
interleavings: every read must see the rows of exactly the day it needs, and nobody may get stuck."""
import random

import pytest


def run_schedule(n_strips, n_days, pick):
    # per strip: mailbox[side][parity] = day tag of the rows it holds, flag[side] = latest slot delivered
    mail = [[[None, None], [None, None]] for _ in range(n_strips)]
    flag = [[0, 0] for _ in range(n_strips)]
    pc = [(0, "wait")] * n_strips          # (day, phase) with phases wait -> read -> push -> signal
    done = [False] * n_strips
    UP, DN = 0, 1                           # side 0: mailbox filled by the strip above, side 1: by the strip below

    def neighbours(s):
        return [(UP, s - 1)] if s == n_strips - 1 and s > 0 else ([(DN, s + 1)] if s == 0 and n_strips > 1 else
                                                                    [(UP, s - 1), (DN, s + 1)] if n_strips > 1 else [])

    def enabled(s):
        if done[s]:
            return False
        x, ph = pc[s]
        if ph == "wait" and x > 0:
            return all(flag[s][side] >= x for side, _ in neighbours(s))
        return True

    steps = 0
    while not all(done):
        ready = [s for s in range(n_strips) if enabled(s)]
        assert ready, "deadlock: %r" % (pc,)
        s = pick(ready, pc)
        x, ph = pc[s]
        if ph == "wait":
            pc[s] = (x, "read")
        elif ph == "read":
            if x > 0:
                for side, _ in neighbours(s):
                    assert mail[s][side][x & 1] == x, "strip %d day %d read rows of day %r" % (s, x, mail[s][side][x & 1])
            pc[s] = (x, "push")
        elif ph == "push":
            for side, nb in neighbours(s):
                mail[nb][1 - side][(x + 1) & 1] = x + 1      # my rows land in the neighbour's mailbox facing me
            pc[s] = (x, "signal")
        else:
            for side, nb in neighbours(s):
                flag[nb][1 - side] = x + 1
            if x + 1 == n_days:
                done[s] = True
            else:
                pc[s] = (x + 1, "wait")
        steps += 1
        assert steps < 10 * n_strips * n_days * 4
    return True


@pytest.mark.parametrize("n_strips", [2, 3, 5, 8])
def test_random_interleavings_never_read_the_wrong_day(n_strips):
    rng = random.Random(n_strips)
    for _ in range(300):
        assert run_schedule(n_strips, 9, lambda ready, pc: rng.choice(ready))


@pytest.mark.parametrize("n_strips", [2, 3, 8])
def test_a_strip_running_as_far_ahead_as_it_can(n_strips):
    for favourite in range(n_strips):
        # always advance the favourite when it can move, otherwise the strip that is furthest behind
        def pick(ready, pc, f=favourite):
            return f if f in ready else min(ready, key=lambda s: pc[s][0])
        assert run_schedule(n_strips, 12, pick)
        # and the opposite: the favourite only moves when nobody else can
        def pick2(ready, pc, f=favourite):
            others = [s for s in ready if s != f]
            return max(others, key=lambda s: pc[s][0]) if others else f
        assert run_schedule(n_strips, 12, pick2)
