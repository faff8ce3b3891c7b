"""CPU model of the scheduler of the compacted / hybrid drain (nesosim_b200/csrc/host_path.inl, the `compact` branch of
nesosim_run_season_host): the main thread's polling loop, transcribed decision by decision, driven by a simulated link
(one FIFO copy queue), a simulated pool of host threads and simulated batch computations with random durations.

What must hold for every interleaving:
  * every (member, array) block of every batch reaches the host exactly once -- packed in a chunk or whole;
  * the link starts writing a ring slot only after the chunk that used the slot before has been scattered;
  * a batch is computed (its device buffers are overwritten) only after every copy out of the batch two before it,
    which used the same buffers, has completed;
  * no copy of a batch starts before the batch is computed and packed;
  * the loop ends.
The same model with the clause `issued < taken + RING` dropped from the issue condition (a slot would count as free
while its previous chunk has not even been handed to the pool) must be caught -- the checks have teeth."""
import random

PLAIN_MAX = 2


class Sim:
    def __init__(self, rng, n_batches, members, n_arr, ring, chunk_target, hybrid, workers, link_rate, scatter_rate,
                 launch_ms, broken=False):
        self.rng = rng
        self.nb, self.ring, self.target, self.hybrid, self.broken = n_batches, ring, chunk_target, hybrid, broken
        self.n_arr = n_arr
        self.blk = [2.0 if i == 0 else 1.0 for i in range(n_arr)]          # packed size of array i (snowDepths is double)
        self.cnt = [members if b < n_batches - 1 else max(1, members - rng.randrange(members)) for b in range(n_batches)]
        self.link_rate, self.scatter_rate, self.launch_ms = link_rate, scatter_rate, launch_ms
        self.worker_free = [0.0] * workers
        self.t = 0.0
        self.link_free = 0.0
        # records
        self.launch_time, self.ready = {}, {}
        self.chunks = []                 # dicts: nb, t0, t1, start, arrive, done (scatter finished; None until taken)
        self.plain = []                  # dicts: nb, t, start, finish
        self.copies_of_batch = {b: [] for b in range(n_batches)}

    # ---- the simulated machine
    def jitter(self, x):
        return x * self.rng.uniform(0.5, 1.5)

    def link_copy(self, nb, size):
        start = max(self.link_free, self.t, self.ready[nb])      # the copy stream waits for the batch's `done` event
        finish = start + self.jitter(size / self.link_rate)
        self.link_free = finish
        self.copies_of_batch[nb].append(finish)
        return start, finish

    def scatter(self, size):
        w = min(range(len(self.worker_free)), key=lambda i: self.worker_free[i])
        start = max(self.t, self.worker_free[w])
        self.worker_free[w] = start + self.jitter(size / self.scatter_rate)
        return self.worker_free[w]

    # ---- host_path.inl, the while loop of the compacted drain
    def run(self):
        nbatch, ring = self.nb, self.ring
        lo = [0] * nbatch
        hi = [0] * nbatch
        last_chunk = [-1] * nbatch
        plain_out = [0] * nbatch
        issued_all = [False] * nbatch
        plain_q = []                                 # FIFO of (finish, nb)
        launched = cur = taken = 0
        steps = 0

        def slot_free(ch):
            if ch < ring:
                return True
            d = self.chunks[ch - ring]["done"]
            if self.broken:                          # pending[] still 0 because the chunk was never taken
                return d is None or d <= self.t
            return d is not None and d <= self.t

        def batch_drained(b):
            return issued_all[b] and taken > last_chunk[b] and plain_out[b] == 0

        while cur < nbatch or taken < len(self.chunks) or plain_q:
            steps += 1
            assert steps < 200000, "the drain loop does not end"
            progress = False
            if launched < nbatch and (launched < 2 or batch_drained(launched - 2)):
                b = launched
                self.launch_time[b] = self.t
                self.t += self.jitter(self.launch_ms)                # run_members blocks the host until the flag is read
                self.ready[b] = self.t + self.jitter(0.3 * self.launch_ms)   # the pack kernel runs behind it
                hi[b] = self.cnt[b] * self.n_arr
                launched += 1
                progress = True
            if cur < launched:
                b = cur
                issued = len(self.chunks)
                if lo[b] == hi[b]:
                    issued_all[b] = True
                    last_chunk[b] = issued - 1
                    cur += 1
                    progress = True
                elif (self.broken or issued < taken + ring) and slot_free(issued):
                    t0 = t1 = lo[b]
                    size = 0.0
                    while True:
                        size += self.blk[t1 % self.n_arr]
                        t1 += 1
                        if not (t1 < hi[b] and size + self.blk[t1 % self.n_arr] <= self.target):
                            break
                    lo[b] = t1
                    start, arrive = self.link_copy(b, size)
                    self.chunks.append({"nb": b, "t0": t0, "t1": t1, "start": start, "arrive": arrive, "done": None})
                    self.t += 0.005
                    progress = True
                elif self.hybrid and len(plain_q) < PLAIN_MAX:
                    hi[b] -= 1
                    t = hi[b]
                    start, finish = self.link_copy(b, 2.2 * self.blk[t % self.n_arr])     # a whole block is ~2.2x its packed size
                    self.plain.append({"nb": b, "t": t, "start": start, "finish": finish})
                    plain_q.append((finish, b))
                    plain_out[b] += 1
                    self.t += 0.005
                    progress = True
            if plain_q and plain_q[0][0] <= self.t:
                plain_out[plain_q.pop(0)[1]] -= 1
                progress = True
            if taken < len(self.chunks) and self.chunks[taken]["arrive"] <= self.t:
                c = self.chunks[taken]
                c["done"] = max(self.scatter(self.blk[t % self.n_arr]) for t in range(c["t0"], c["t1"]))
                taken += 1
                self.t += 0.002
                progress = True
            if not progress:                         # sleep; jump to the next thing that can change what the loop sees
                nxt = []
                if plain_q:
                    nxt.append(plain_q[0][0])
                if taken < len(self.chunks):
                    nxt.append(self.chunks[taken]["arrive"])
                nxt += [c["done"] for c in self.chunks[max(0, len(self.chunks) - ring):] if c["done"] is not None]
                nxt = [x for x in nxt if x > self.t]
                self.t = max(self.t + 0.01, min(nxt)) if nxt else self.t + 0.01
        return self

    # ---- what must hold
    def check(self):
        for b in range(self.nb):
            want = set(range(self.cnt[b] * self.n_arr))
            got = [t for c in self.chunks if c["nb"] == b for t in range(c["t0"], c["t1"])] + [p["t"] for p in self.plain if p["nb"] == b]
            assert sorted(got) == sorted(want), "batch %d: blocks delivered %r" % (b, sorted(got))
        for i, c in enumerate(self.chunks):
            assert c["done"] is not None and c["done"] >= c["arrive"] >= c["start"] >= self.ready[c["nb"]]
            if i >= self.ring:
                prev = self.chunks[i - self.ring]
                assert prev["done"] is not None and prev["done"] <= c["start"], \
                    "ring slot %d overwritten at %.3f while chunk %d is scattered until %s" % (i % self.ring, c["start"], i - self.ring, prev["done"])
        for p in self.plain:
            assert p["start"] >= self.ready[p["nb"]]
        for b in range(2, self.nb):
            assert self.launch_time[b] >= max(self.copies_of_batch[b - 2]), "batch %d computed over buffers still being copied" % b


def _random_sim(seed, **force):
    rng = random.Random(seed)
    kw = dict(n_batches=rng.randint(1, 5), members=rng.randint(1, 6), n_arr=rng.choice([1, 2, 9]), ring=rng.randint(2, 6),
              chunk_target=rng.choice([0.5, 1.0, 3.0, 8.0]), hybrid=rng.random() < 0.7, workers=rng.choice([1, 2, 4, 16]),
              link_rate=rng.choice([0.2, 1.0, 5.0, 50.0]), scatter_rate=rng.choice([0.05, 0.5, 2.0, 20.0]),
              launch_ms=rng.choice([0.01, 0.5, 3.0]))
    kw.update(force)
    return Sim(rng, **kw)


def test_every_interleaving_delivers_every_block_once_and_respects_the_buffers():
    for seed in range(600):
        _random_sim(seed).run().check()


def test_the_split_follows_the_bottleneck():
    """Slow host threads, fast link: whole blocks go over the link.  Hybrid off: none do."""
    slow_cpu = _random_sim(1, n_batches=3, members=6, n_arr=9, ring=3, chunk_target=3.0, hybrid=True, workers=1,
                           link_rate=50.0, scatter_rate=0.05, launch_ms=0.01).run()
    slow_cpu.check()
    n_blocks = sum(slow_cpu.cnt) * 9
    assert len(slow_cpu.plain) > n_blocks // 2
    off = _random_sim(1, n_batches=3, members=6, n_arr=9, ring=3, chunk_target=3.0, hybrid=False, workers=1,
                      link_rate=50.0, scatter_rate=0.05, launch_ms=0.01).run()
    off.check()
    assert not off.plain
    # fast threads, slow link: (almost) everything packed -- the ring holds queued copies, so a few blocks still go whole
    slow_link = _random_sim(1, n_batches=3, members=6, n_arr=9, ring=6, chunk_target=3.0, hybrid=True, workers=16,
                            link_rate=0.2, scatter_rate=20.0, launch_ms=0.01).run()
    slow_link.check()
    assert len(slow_link.plain) < n_blocks // 4


def test_a_slot_counted_free_before_its_chunk_was_taken_is_caught():
    failures = 0
    for seed in range(200):
        try:
            _random_sim(seed, broken=True).run().check()
        except AssertionError:
            failures += 1
    assert failures > 20
