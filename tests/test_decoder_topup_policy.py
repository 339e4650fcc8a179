"""Host-side model of the decoders' word top-up (csrc/ans_kernels.cuh `top_up`, also range_kernels.cuh / chain_kernels.cuh):
a lane's 16-word ring is refilled by 16-byte blocks, one request per check (every 4 symbols), two blocks in flight
(`cp.async.wait_group 1`: a block is only guaranteed to have landed at the second check after its request).

The coder treats an empty ring as "stream exhausted" (stack.rs:1091-1097: no refill when there is no word), so the
policy must never let the ring run empty while words of the stream are still unstaged or in flight -- whatever the
symbols are (a lane pops at most one word per symbol).  This test replays the kernel's bookkeeping step by step for
every stream length and alignment of the first block under the worst case (a pop at every symbol) and under random
pop patterns, with blocks landing as late as the kernel allows.  It also checks that the ring never holds more than
its 16 words (a landing block must not overwrite unread words)."""
import numpy as np
import pytest

RING = 16      # kDecRingWords
CHECK = 4      # kCheckEvery


class Lane:
    def __init__(self, total_words, first_block_words):
        self.unstaged = total_words
        self.avail = 0            # landed and unread
        self.pending = 0          # words of the newest request (in flight)
        self.pending_old = 0      # words of the request before it (in flight until the next check)
        self.first = first_block_words

    def request(self, block_words):
        n = min(self.unstaged, block_words)
        self.unstaged -= n
        self.pending = n
        assert self.avail + self.pending + self.pending_old <= RING, "a landing block would overwrite unread words"

    def top_up(self):
        self.avail += self.pending_old          # cp.async.wait_group 1: everything but the newest group has landed
        self.pending_old, self.pending = self.pending, 0
        if self.avail + self.pending_old <= RING - 4 and self.unstaged != 0:
            self.request(4)

    def start(self):
        if self.unstaged:
            self.request(self.first)
        for _ in range(3):
            self.top_up()
        self.avail += self.pending + self.pending_old   # cp.async.wait_group 0
        self.pending = self.pending_old = 0

    def left(self):
        return self.unstaged + self.pending + self.pending_old + self.avail

    def pop(self):
        """one refill attempt of the coder; returns False if the ring looks empty"""
        if self.avail == 0:
            assert self.unstaged + self.pending + self.pending_old == 0, "ring empty while the stream still has words"
            return False
        self.avail -= 1
        return True


def run(total_words, first, pops):
    lane = Lane(total_words, first)
    lane.start()
    popped = 0
    for _ in range(2):           # initial state: up to two words
        popped += lane.pop()
    lane.top_up()
    it = iter(pops)
    while lane.left():
        for _ in range(CHECK):
            if next(it):
                popped += lane.pop()
        lane.top_up()
    assert popped == total_words


@pytest.mark.parametrize("first", [1, 2, 3, 4])
def test_ring_never_runs_dry_when_every_symbol_pops(first):
    for total in list(range(0, 200)) + [1000, 4097]:
        run(total, first, iter(lambda: True, None))


def test_ring_never_runs_dry_under_random_pop_patterns():
    rng = np.random.default_rng(5)
    for _ in range(300):
        total = int(rng.integers(0, 600))
        p = float(rng.choice([0.05, 0.17, 0.5, 0.9, 1.0]))
        pops = iter(lambda: bool(rng.random() < p), None)
        # bursts: a few checks with a pop at every symbol in between
        def bursty():
            while True:
                if rng.random() < 0.1:
                    for _ in range(int(rng.integers(4, 40))):
                        yield True
                yield next(pops)
        run(total, int(rng.integers(1, 5)), bursty())
