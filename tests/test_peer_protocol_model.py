"""
CPU model of the peer-memory exchange protocol (csrc/peer.cu): R ranks, each with a double-buffered data
arena and a flag block; per epoch e a rank (1) writes its data of epoch e into buffer e & 1, (2) stores e into its slot of every
peer's flag block, (3) waits until its own flag block shows >= e from every peer, (4) pulls buffer e & 1 from every peer.
The claim the CUDA code relies on: two buffers suffice -- a rank never pulls data of another epoch, whatever the interleaving.
The model runs the ranks as generators under a seeded random scheduler (every interleaving of the steps is reachable) and also
checks that the claim is FALSE with a single buffer, i.e. that the test can see the race it guards against.
"""
import random

import pytest


def rank_program(r, R, epochs, n_buf, data, flags, log):
    """One rank's steps; yields between every shared-memory access so the scheduler can interleave anything."""
    for e in range(1, epochs + 1):
        data[r][e % n_buf] = (r, e)                                   # (1) the step's kernels write this epoch's buffer
        yield
        for q in range(R):                                            # (2) signal: st.release.sys into every flag block
            if q != r:
                flags[q][r] = e
                yield
        for q in range(R):                                            # (3) wait: spin on the LOCAL flag block
            while q != r and flags[r][q] < e:
                yield
        for q in range(R):                                            # (4) pull from the peers
            if q != r:
                log.append((r, e, q, data[q][e % n_buf]))
                yield


def run(R, epochs, n_buf, seed):
    rng = random.Random(seed)
    data = [[None] * n_buf for _ in range(R)]
    flags = [[0] * R for _ in range(R)]
    log = []
    live = [rank_program(r, R, epochs, n_buf, data, flags, log) for r in range(R)]
    steps = 0
    while live:
        g = rng.choice(live) if rng.random() < 0.7 else live[0]      # biased: lets one rank run far ahead of the others
        try:
            next(g)
        except StopIteration:
            live.remove(g)
        steps += 1
        assert steps < 5_000_000, 'model deadlocked'
    return log


@pytest.mark.parametrize('R', [2, 3, 8])
def test_double_buffering_is_sufficient(R):
    for seed in range(40):
        for (r, e, q, seen) in run(R, epochs=12, n_buf=2, seed=seed):
            assert seen == (q, e), f'rank {r} pulled {seen} from rank {q} at epoch {e} (seed {seed})'


def test_single_buffer_races_and_the_model_sees_it():
    bad = 0
    for seed in range(40):
        bad += sum(1 for (r, e, q, seen) in run(3, epochs=12, n_buf=1, seed=seed) if seen != (q, e))
    assert bad > 0


def test_no_rank_runs_more_than_one_epoch_ahead():
    """The reason two buffers suffice: rank r can finish epoch e only after every peer has signalled e, so no peer is behind e."""
    R, epochs = 4, 10
    for seed in range(20):
        first_pull = {}
        for (r, e, q, seen) in run(R, epochs, 2, seed):
            first_pull.setdefault((r, e), len(first_pull))
        order = sorted(first_pull, key=first_pull.get)                # (rank, epoch) in the order their pulls began
        started = {}
        for (r, e) in order:
            started[r] = e
            assert max(started.values()) - min(started.get(x, 0) for x in range(R)) <= 1 or len(started) < R
