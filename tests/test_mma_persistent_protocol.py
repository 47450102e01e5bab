"""Host model of the barrier protocol of hamming_mma_persistent_kernel (csrc/hamming_mma_persistent.cuh — the
persistent form of the tcgen05 matcher: built, proven bit-identical on the GPU, measured and not shipped; compiled in
development builds only).  The model is kept because it is what found the kernel's first bug.

The persistent tcgen05 matcher keeps every mbarrier running across the jobs of a CTA; each role (job fetch +
TMA lane, expander threads, MMA lane, the two epilogue groups) counts phases on its own.  An mbarrier wait only
sees the PARITY of a phase, so a role that waits for phase k while the barrier is still in phase k - 1 passes
at once (it mistakes phase k - 2 for k), and one that is two phases late waits for a phase that may never come.
This test restates the kernel's wait / arrive sequence role by role (same counters, same expressions) and runs
it under random interleavings over random job mixes (one- and two-tile jobs, stages the bulk copy cannot take):
no deadlock, and every wait that passes passes for the phase it meant.  It is a model of the protocol, not of the
arithmetic — the keys are checked on the GPU (tests/test_gpu_matching.py under both matcher kernels).
"""
import random

import pytest

NB, JR = 2, 4


class Barrier:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"{self.name}: more arrivals than the barrier expects"
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def passes(self, parity):   # mbarrier.try_wait.parity
        return (self.phase & 1) != parity


class Wait:
    def __init__(self, bar, phase):   # wait for phase `phase` (0-based) of `bar` to complete
        self.bar, self.phase = bar, phase


def acc_uses(job, acc):
    n_stage, n_tiles = job["n_stage"], job["n_tiles"]
    if n_tiles == 2:
        return n_stage
    return (n_stage + 1) >> 1 if acc == 0 else n_stage >> 1


def acc_of(job, s, t):
    return t if job["n_tiles"] == 2 else ((job["n_stage"] - 1 - s) & 1)


class Model:
    def __init__(self, jobs, n_exp=2, n_epi=2, seed=0):
        self.jobs = jobs + [None]                     # the sentinel the TMA lane publishes last
        self.rng = random.Random(seed)
        self.n_exp, self.n_epi = n_exp, n_epi
        readers = 2 * n_epi + n_exp + 1
        B = Barrier
        self.raw_full = [B(f"raw_full{i}", 1) for i in range(NB)]
        self.raw_empty = [B(f"raw_empty{i}", n_exp) for i in range(NB)]
        self.b_full = [B(f"b_full{i}", n_exp) for i in range(NB)]
        self.b_empty = [B(f"b_empty{i}", 1) for i in range(NB)]
        self.d_full = [B(f"d_full{i}", 1) for i in range(2)]
        self.d_empty = [B(f"d_empty{i}", n_epi) for i in range(2)]
        self.a_ready = [B(f"a_ready{i}", n_epi) for i in range(2)]
        self.sched_full = [B(f"sched_full{i}", 1) for i in range(JR)]
        self.sched_empty = [B(f"sched_empty{i}", readers) for i in range(JR)]
        self.ring = [None] * JR
        self.commits = []          # tcgen05.commit arrivals still in flight, in issue order
        self.mma_log = []          # (job, stage, tile, accumulator) in issue order
        self.fold_log = []         # (group, job, stage)
        self.a_tile = [None, None] # job whose +-1 tile sits in TMEM, per tile
        self.mma_done = 0          # MMA jobs whose commit has landed
        self.check_tiles = True

    # ---- roles -------------------------------------------------------------------------------------
    def read_job(self, n):
        yield Wait(self.sched_full[n % JR], n // JR)
        job = self.ring[n % JR]
        assert job is None or job["n"] == n, "ring slot overwritten before it was read"
        self.sched_empty[n % JR].arrive()
        return job

    def tma_lane(self):
        def fetch(n):
            if n >= JR:
                yield Wait(self.sched_empty[n % JR], n // JR - 1)
            self.ring[n % JR] = self.jobs[n]
            self.sched_full[n % JR].arrive()
            return self.jobs[n]
        gsb, armed = 0, [0, 0]
        cur = yield from fetch(0)
        n = 0
        while cur is not None:
            nxt = None
            fetch_at = min(cur["n_stage"], NB) - 1
            for s in range(cur["n_stage"]):
                b = (gsb + s) % NB
                if cur["tma"][s]:
                    if armed[b] > 0:
                        yield Wait(self.raw_empty[b], armed[b] - 1)
                    armed[b] += 1
                    self.raw_full[b].arrive()     # arrive.expect_tx + the copy landing, as one event
                if s == fetch_at:
                    nxt = yield from fetch(n + 1)
            gsb += cur["n_stage"]
            cur = nxt
            n += 1

    def expander(self):
        gsb, raw_cnt, n = 0, [0, 0], 0
        while True:
            cur = yield from self.read_job(n)
            if cur is None:
                return
            for s in range(cur["n_stage"]):
                g = gsb + s
                b = g % NB
                if cur["tma"][s]:
                    yield Wait(self.raw_full[b], raw_cnt[b])
                    raw_cnt[b] += 1
                if g >= NB:
                    yield Wait(self.b_empty[b], g // NB - 1)
                if cur["tma"][s]:
                    self.raw_empty[b].arrive()
                self.b_full[b].arrive()
            gsb += cur["n_stage"]
            n += 1

    def mma_lane(self):
        st = dict(gsb=0, acc_base=[0, 0], a_cnt=[0, 0])

        def wait_mma(job, s, t):
            if s == 0:
                yield Wait(self.a_ready[t], st["a_cnt"][t])
                st["a_cnt"][t] += 1
            g = st["gsb"] + s
            if t == 0:
                yield Wait(self.b_full[g % NB], g // NB)
            acc = acc_of(job, s, t)
            use = st["acc_base"][acc] + (s if job["n_tiles"] == 2 else s >> 1)
            if use > 0:
                yield Wait(self.d_empty[acc], use - 1)

        n = 0
        cur = yield from self.read_job(0)
        s = t = 0
        if cur is not None:
            yield from wait_mma(cur, 0, 0)
        while cur is not None:
            g = st["gsb"] + s
            b, acc = g % NB, acc_of(cur, s, t)
            ns, nt = s, t + 1
            if nt == cur["n_tiles"]:
                nt, ns = 0, s + 1
            last = ns == cur["n_stage"]
            assert not self.check_tiles or self.a_tile[t] == cur["n"], (
                f"MMA of job {cur['n']} reads the +-1 tile of job {self.a_tile[t]}")
            self.mma_log.append((cur["n"], s, t, acc))
            if not last:
                yield from wait_mma(cur, ns, nt)
            arrivals = [self.d_full[acc]]
            if t == cur["n_tiles"] - 1:
                arrivals.insert(0, self.b_empty[b])
            self.commits.append((arrivals, (cur["n"], t)))
            if last:
                st["gsb"] += cur["n_stage"]
                st["acc_base"][0] += acc_uses(cur, 0)
                st["acc_base"][1] += acc_uses(cur, 1)
                n += 1
                cur = yield from self.read_job(n)
                s = t = 0
                if cur is not None:
                    yield from wait_mma(cur, 0, 0)
            else:
                s, t = ns, nt

    def tensor_pipe(self):   # commits land later, in order
        if not self.commits:
            return False
        arrivals, tag = self.commits.pop(0)
        for bar in arrivals:
            bar.arrive()
        self.mma_done += 1
        self.last_done = getattr(self, "last_done", {})
        self.last_done[tag] = self.last_done.get(tag, 0) + 1
        return True

    def epilogue(self, tile):
        def build_a(job_n):
            # every MMA that read the old tile must be complete: none of its MMAs may sit in the commit queue
            old = self.a_tile[tile]
            assert all(tag != (old, tile) for _, tag in self.commits), "+-1 tile rewritten under running MMAs"
            issued = [m for m in self.mma_log if m[0] == old and m[2] == tile]
            want = 0 if old is None else next(j for j in self.jobs if j and j["n"] == old)["n_stage"]
            if old is not None and tile < next(j for j in self.jobs if j and j["n"] == old)["n_tiles"]:
                assert len(issued) == want, "+-1 tile rewritten before its job issued all MMAs"
            self.a_tile[tile] = job_n
            self.a_ready[tile].arrive()

        def peek(n):
            yield Wait(self.sched_full[n % JR], n // JR)
            job = self.ring[n % JR]
            assert job is None or job["n"] == n
            return job is not None and tile < job["n_tiles"]

        acc_base, prebuilt, n = 0, False, 0
        yield from peek(0)
        while True:
            cur = yield from self.read_job(n)
            if cur is None:
                return
            if tile < cur["n_tiles"] and not prebuilt:
                build_a(cur["n"])
            prebuilt = False
            nxt_builds = yield from peek(n + 1)
            two = cur["n_tiles"] == 2
            my_n = acc_uses(cur, tile)
            first, step = (0, 1) if two else ((cur["n_stage"] - 1 - tile) & 1, 2)
            for i in range(my_n):
                s = first + i * step
                use = acc_base + i
                yield Wait(self.d_full[tile], use)
                assert acc_of(cur, s, tile if two else 0) == tile
                self.fold_log.append((tile, cur["n"], s))
                self.d_empty[tile].arrive()
                if i == my_n - 1 and nxt_builds:
                    build_a(cur["n"] + 1)
                    prebuilt = True
            acc_base += my_n
            n += 1

    # ---- scheduler ---------------------------------------------------------------------------------
    def run(self):
        # the epilogue agents of a group each arrive; only the first of them moves the modelled TMEM tile
        agents = {"tma": self.tma_lane(), "mma": self.mma_lane()}
        for i in range(self.n_exp):
            agents[f"exp{i}"] = self.expander()
        for tile in range(2):
            for i in range(self.n_epi):
                agents[f"epi{tile}.{i}"] = self.epilogue(tile) if i == 0 else self.epilogue_follower(tile)
        blocked = {}
        for k, gen in list(agents.items()):
            blocked[k] = self.step(gen, None)
            if blocked[k] == "done":
                del agents[k], blocked[k]
        steps = 0
        while agents:
            steps += 1
            assert steps < 2_000_000, "livelock"
            ready = [k for k in agents if self.ready(blocked[k])]
            pipe = bool(self.commits)
            if not ready and not pipe:
                state = {k: (w.bar.name, w.phase, w.bar.phase) for k, w in blocked.items()}
                raise AssertionError(f"deadlock: {state}")
            if pipe and (not ready or self.rng.random() < 0.3):
                self.tensor_pipe()
                continue
            k = self.rng.choice(ready)
            w = blocked[k]
            # the wait passes: it must pass for the phase it meant
            assert w.bar.phase == w.phase + 1, (
                f"{k}: wait for phase {w.phase} of {w.bar.name} passed while the barrier is in phase {w.bar.phase}")
            blocked[k] = self.step(agents[k], None)
            if blocked[k] == "done":
                del agents[k], blocked[k]

    def epilogue_follower(self, tile):
        """the other threads of an epilogue group: same waits and arrivals, no modelled TMEM state"""
        acc_base, prebuilt, n = 0, False, 0
        yield Wait(self.sched_full[0], 0)
        while True:
            cur = yield from self.read_job(n)
            if cur is None:
                return
            if tile < cur["n_tiles"] and not prebuilt:
                self.a_ready[tile].arrive()
            prebuilt = False
            yield Wait(self.sched_full[(n + 1) % JR], (n + 1) // JR)
            nj = self.ring[(n + 1) % JR]
            nxt_builds = nj is not None and tile < nj["n_tiles"]
            my_n = acc_uses(cur, tile)
            for i in range(my_n):
                yield Wait(self.d_full[tile], acc_base + i)
                self.d_empty[tile].arrive()
                if i == my_n - 1 and nxt_builds:
                    self.a_ready[tile].arrive()
                    prebuilt = True
            acc_base += my_n
            n += 1

    @staticmethod
    def step(gen, _):
        try:
            return next(gen)
        except StopIteration:
            return "done"

    @staticmethod
    def ready(w):
        return w.bar.passes(w.phase & 1)


def make_jobs(rng, n_jobs, max_stage, p_one_tile, p_no_tma):
    jobs = []
    for n in range(n_jobs):
        n_stage = rng.randint(1, max_stage)
        whole_job_unaligned = rng.random() < p_no_tma
        tma = [not whole_job_unaligned for _ in range(n_stage)]
        if rng.random() < 0.3:
            tma[-1] = False      # a last stage with fewer rows than the bulk copy's quantum
        jobs.append(dict(n=n, n_stage=n_stage, n_tiles=1 if rng.random() < p_one_tile else 2, tma=tma))
    return jobs


@pytest.mark.parametrize("seed", range(40))
def test_protocol_random_job_mixes(seed):
    rng = random.Random(1000 + seed)
    jobs = make_jobs(rng, rng.randint(0, 14), rng.choice([1, 2, 3, 7]), rng.choice([0.0, 0.3, 1.0]),
                     rng.choice([0.0, 0.5, 1.0]))
    m = Model(jobs, seed=seed)
    m.run()
    # every MMA job issued once, in order, and every accumulator use folded by the group that owns it
    want = [(j["n"], s, t) for j in jobs for s in range(j["n_stage"]) for t in range(j["n_tiles"])]
    assert [x[:3] for x in m.mma_log] == want
    folds = sorted((n, s, g) for g, n, s in m.fold_log)
    want_folds = sorted((j["n"], s, acc_of(j, s, t)) for j in jobs for s in range(j["n_stage"]) for t in range(j["n_tiles"]))
    assert folds == want_folds


def test_protocol_detects_a_group_that_skips_phases():
    """The first version of the kernel let tile 1's warps sit out one-tile jobs while tile 0's warps drained both
    accumulators; the model must catch that (a wait passing for the wrong phase, or a deadlock)."""
    class Broken(Model):
        def epilogue_follower(self, tile):
            acc_base, n = [0, 0], 0
            while True:
                cur = yield from self.read_job(n)
                if cur is None:
                    return
                if tile < cur["n_tiles"]:
                    self.a_ready[tile].arrive()
                    for s in range(cur["n_stage"]):
                        acc = tile if cur["n_tiles"] == 2 else (s & 1)
                        use = acc_base[acc] + (s if cur["n_tiles"] == 2 else s >> 1)
                        yield Wait(self.d_full[acc], use)
                        self.d_empty[acc].arrive()
                acc_base[0] += acc_uses(cur, 0)
                acc_base[1] += acc_uses(cur, 1)
                n += 1
        epilogue = epilogue_follower
    jobs = [dict(n=0, n_stage=3, n_tiles=2, tma=[True] * 3), dict(n=1, n_stage=3, n_tiles=1, tma=[True] * 3),
            dict(n=2, n_stage=4, n_tiles=2, tma=[True] * 4)]
    caught = 0
    for seed in range(20):
        m = Broken(jobs, seed=seed)
        m.check_tiles = False
        try:
            m.run()
        except AssertionError:
            caught += 1
    assert caught > 0
