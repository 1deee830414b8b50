"""Model check of the halo mailbox flow control of csrc/sphk_mg.cu (no GPU, no library calls): ranks on a line, every
halo kernel k (1) stores its message into the neighbours' mailbox k & 1 and raises their flag to k, (2) waits until its
own flags reach k, (3) reads its own mailboxes k & 1.  Kernels of one rank run in stream order; ranks interleave
arbitrarily.  Claim under test: two mailboxes per direction suffice -- a mailbox is never overwritten before it was
read and never read before it was written -- and the protocol cannot deadlock."""
import random

import pytest
from hypothesis import given, settings, strategies as st


class Rank:
    def __init__(self, r, world):
        self.r, self.world = r, world
        self.k = 1                  # sequence number of the running halo kernel
        self.phase = 0              # 0: send, 1: wait, 2: read
        self.flag = {-1: 0, +1: 0}  # last sequence published by the neighbour on that side
        self.box = {(-1, 0): None, (-1, 1): None, (+1, 0): None, (+1, 1): None}   # (side, parity) -> (sender, seq)
        self.unread = set()         # mailboxes holding a message that has not been consumed yet

    def sides(self):
        return [d for d in (-1, +1) if 0 <= self.r + d < self.world]


def run(world, halos, schedule, nbox=2):
    ranks = [Rank(r, world) for r in range(world)]
    picks = iter(schedule)
    idle = 0
    while any(x.k <= halos for x in ranks):
        try:
            r = next(picks) % world
        except StopIteration:
            r = random.randrange(world)
        x = ranks[r]
        if x.k > halos:
            continue
        if x.phase == 0:                                    # remote stores + publish
            for d in x.sides():
                y = ranks[r + d]
                key = (-d, x.k % nbox)                      # I am the neighbour's side -d
                assert key not in y.unread, f"rank {r} overwrote an unread mailbox of rank {r + d} (message {x.k})"
                y.box[key] = (r, x.k)
                y.unread.add(key)
                y.flag[-d] = x.k
            x.phase, idle = 1, 0
        elif x.phase == 1:                                  # wait on the local flags
            if all(x.flag[d] >= x.k for d in x.sides()):
                x.phase, idle = 2, 0
            else:
                idle += 1
                assert idle < 10000 * world, "deadlock: no rank can make progress"
        else:                                               # unpack
            for d in x.sides():
                key = (d, x.k % nbox)
                assert x.box[key] == (r + d, x.k), f"rank {r} read {x.box[key]} instead of message {x.k} from rank {r + d}"
                x.unread.discard(key)
            x.k, x.phase, idle = x.k + 1, 0, 0
    return True


@settings(max_examples=200, deadline=None)
@given(world=st.integers(2, 5), halos=st.integers(1, 12), schedule=st.lists(st.integers(0, 4), max_size=400))
def test_two_mailboxes_per_direction_suffice(world, halos, schedule):
    random.seed(len(schedule) * 131 + world)
    assert run(world, halos, schedule)


def test_single_mailbox_would_not_suffice():
    """The checker is not vacuous: with ONE mailbox per direction a fast neighbour overwrites an unread message
    (rank 0 sends 1, rank 1 sends 1, rank 0 reads 1 and sends 2 before rank 1 has read message 1)."""
    with pytest.raises(AssertionError, match="overwrote an unread mailbox"):
        run(2, 3, [0, 1, 0, 0, 0], nbox=1)
