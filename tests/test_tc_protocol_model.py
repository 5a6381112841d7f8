"""Executable model of the synchronisation protocol of conv_fwd_tc_kernel (btcdet_b200/csrc/sparse_conv_tc.cu).

The kernel's roles — index loader / tile scheduler, G producer groups, MMA issuer (two stages per trip), weight loader,
epilogue — are restated as cooperative agents over modelled mbarriers (arrival counts + phase parity), an in-order tensor
pipe that retires MMA stages and tcgen05.commit arrivals asynchronously, and asynchronous TMA completions.  A random
scheduler explores interleavings; the model asserts what the hardware run can only show as a hang or a wrong number:

  * no deadlock (every CTA reaches its final barrier), with variable chunk counts per tile, including runs of 1-chunk
    tiles (producer groups that own no stage in consecutive tiles) and fewer tiles than CTAs;
  * every MMA stage reads the A slot and the weight slot that were filled for exactly that (tile, chunk);
  * no ring slot / index buffer / accumulator is overwritten before its last reader is done;
  * every tile is processed exactly once and the epilogue sees exactly its chunk list;
  * the dynamic scheduler's counter is back at zero after the launch.

Parameters cover the shipped configuration (STAGES 4 / 6, G 4 / 2, one commit per stage) and, as a property of the model
only, commit groups (CG 2 / 3: measured on hardware at the start of round 2 — slower — and removed from the kernel).
Round 2: the weight tile's arrival shares the stage's `full` barrier with the producer group (one wait per stage in the
issuer instead of two).  Keep in sync with the kernel by hand."""
import random

import pytest


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.phase ^= 1
            self.pending = self.count

    def ready(self, parity):          # mbarrier.try_wait.parity: the phase with this parity has completed
        return self.phase != parity


class Blocked(Exception):
    pass


def wait(bar, parity):
    if not bar.ready(parity):
        raise Blocked()


class CTA:
    def __init__(self, cid, sim):
        self.cid, self.sim = cid, sim
        p = sim.p
        S, G = p["STAGES"], p["G"]
        # one barrier per stage: the producer group (128 threads in the kernel) + the weight loader (arrive + tx bytes)
        self.full = [MBar(2) for _ in range(S)]
        self.empty = [MBar(1) for _ in range(S // p["CG"])]
        self.tmem_full = [MBar(2), MBar(2)]
        self.tmem_empty = [MBar(1), MBar(1)]               # the epilogue (128 threads in the kernel)
        self.nbr_empty = [MBar(G + 2), MBar(G + 2)]
        self.list_full = [MBar(1), MBar(1)]
        self.s_tile, self.s_cnt, self.s_list, self.s_epi = [None, None], [0, 0], [None, None], [None, None]
        self.buf_readers = [set(), set()]                  # who still reads index buffer b
        self.a_slot = [None] * S                           # (tile, chunk) held, None = free / consumed
        self.b_slot = [None] * S
        self.acc = [None, None]                            # chunk lists accumulated per accumulator buffer
        self.acc_busy = [False, False]
        self.pipe = []                                     # in-order tensor pipe: ("mma", ...) / ("commit", bar)
        self.async_ev = []                                 # pending TMA completions: callables
        self.done_roles = set()
        self.agents = {"idx": self.idx(), "mma": self.mma(), "bld": self.bld(), "epi": self.epi()}
        for g in range(G):
            self.agents["prod%d" % g] = self.producer(g)

    # ---- roles (generators: yield = scheduling point, Blocked = retry later) -------------------------------------
    def idx(self):
        sim, p = self.sim, self.sim.p
        tl = 0
        while True:
            buf = tl & 1
            while True:
                try:
                    wait(self.nbr_empty[buf], ((tl >> 1) & 1) ^ 1)
                    break
                except Blocked:
                    yield "blocked"
            assert not self.buf_readers[buf], ("index buffer overwritten while in use", self.buf_readers[buf])
            if not p["DYN"]:
                my = (sim.num_tiles - self.cid + sim.grid - 1) // sim.grid
                tile = self.cid + tl * sim.grid if tl < my else -1
            elif tl == 0:
                tile = self.cid
            else:
                t = sim.counter
                sim.counter += 1
                sim.fetches += 1
                tile = sim.P + t
                if tile >= sim.num_tiles:
                    tile = -1
                    if t == sim.num_tiles - 1:
                        sim.counter = 0
            if tile < 0:
                self.s_tile[buf], self.s_cnt[buf] = -1, 0
                self.list_full[buf].arrive()
                return
            yield "tma idx"                                 # TMA load + mask computation
            self.s_tile[buf], self.s_list[buf] = tile, list(sim.tiles[tile])
            self.s_cnt[buf] = len(sim.tiles[tile])
            self.buf_readers[buf] = set(["mma", "bld"] + ["prod%d" % g for g in range(p["G"])])
            self.list_full[buf].arrive()
            tl += 1
            yield "next"

    def producer(self, g):
        p = self.sim.p
        S, G, CG, DEPTH = p["STAGES"], p["G"], p["CG"], p["DEPTH"]
        me = "prod%d" % g
        st = dict(it_tile=0, it_pos=g, cur_tile=-1, cur_cnt=0, done=False)
        inflight = []
        s, ph = g % S, 0

        def locate(may_block):
            while True:
                if st["cur_tile"] != st["it_tile"]:
                    buf = st["it_tile"] & 1
                    par = (st["it_tile"] >> 1) & 1
                    if not self.list_full[buf].ready(par):
                        if may_block:
                            raise Blocked()
                        return False
                    if self.s_tile[buf] < 0:
                        st["done"] = True
                        return False
                    st["cur_tile"], st["cur_cnt"] = st["it_tile"], self.s_cnt[buf]
                if st["it_pos"] < st["cur_cnt"]:
                    return True
                st["it_pos"] -= st["cur_cnt"]
                st["it_tile"] += 1
                b = st["cur_tile"] & 1
                self.buf_readers[b].discard(me)
                self.nbr_empty[b].arrive()
                st["cur_tile"] = -1

        while True:
            while not st["done"] and len(inflight) < DEPTH:
                try:
                    ok = locate(bool(p.get("ALWAYS_BLOCK")) or len(inflight) == 0)
                except Blocked:
                    yield "blocked"
                    continue
                if not ok:
                    break
                buf = st["it_tile"] & 1
                assert me in self.buf_readers[buf]
                inflight.append((self.s_tile[buf], self.s_list[buf][st["it_pos"]]))     # issue(): gather in flight
                st["it_pos"] += G
                yield "issued"
            if not inflight:
                return
            item = inflight.pop(0)
            while True:
                try:
                    wait(self.empty[s // CG], ph ^ 1)
                    break
                except Blocked:
                    yield "blocked"
            assert self.a_slot[s] is None, ("A slot overwritten before its MMAs retired", s, self.a_slot[s])
            self.a_slot[s] = item                            # tcgen05.st + wait::st
            self.full[s].arrive()
            s += G
            if s >= S:
                s -= S
                ph ^= 1
            yield "stage"

    def bld(self):
        p = self.sim.p
        S, CG = p["STAGES"], p["CG"]
        sb, pb, tl = 0, 0, 0
        while True:
            buf = tl & 1
            while True:
                try:
                    wait(self.list_full[buf], (tl >> 1) & 1)
                    break
                except Blocked:
                    yield "blocked"
            if self.s_tile[buf] < 0:
                return
            tile, cnt = self.s_tile[buf], self.s_cnt[buf]
            for j in range(cnt):
                chunk = self.s_list[buf][j]
                while True:
                    try:
                        wait(self.empty[sb // CG], pb ^ 1)
                        break
                    except Blocked:
                        yield "blocked"
                assert self.b_slot[sb] is None, ("weight slot overwritten before its MMAs retired", sb)
                slot, tag = sb, (tile, chunk)

                def land(slot=slot, tag=tag):
                    self.b_slot[slot] = tag
                    self.full[slot].arrive()
                self.b_slot[sb] = "in flight"
                self.async_ev.append(land)
                sb += 1
                if sb == S:
                    sb, pb = 0, pb ^ 1
                yield "tma w"
            self.buf_readers[buf].discard("bld")
            self.nbr_empty[buf].arrive()
            tl += 1

    def mma(self):
        p = self.sim.p
        S, CG = p["STAGES"], p["CG"]
        s, ph, tl = 0, 0, 0

        def issue_stage(sa, tile, chunk, buf, first):
            self.pipe.append(("mma", sa, (tile, chunk), buf, first))
            if CG == 1 or sa % CG == CG - 1:
                self.pipe.append(("commit", self.empty[sa // CG], list(range(sa - sa % CG, sa + 1))))

        while True:
            buf = tl & 1
            while True:
                try:
                    wait(self.list_full[buf], (tl >> 1) & 1)
                    break
                except Blocked:
                    yield "blocked"
            tile, cnt = self.s_tile[buf], self.s_cnt[buf]
            chunks = list(self.s_list[buf]) if tile >= 0 else []
            while True:
                try:
                    wait(self.tmem_empty[buf], ((tl >> 1) & 1) ^ 1)
                    break
                except Blocked:
                    yield "blocked"
            assert not self.acc_busy[buf], "accumulator reused before the epilogue drained it"
            self.s_epi[buf] = tile
            self.tmem_full[buf].arrive()
            if tile < 0:
                self.tmem_full[buf].arrive()
                return
            self.buf_readers[buf].discard("mma")
            self.nbr_empty[buf].arrive()
            self.acc_busy[buf] = True
            j = 0
            while j + 1 < cnt:                               # two stages per trip
                s1, ph1 = s + 1, ph
                if s1 == S:
                    s1, ph1 = 0, ph1 ^ 1
                for bar, par in ((self.full[s], ph), (self.full[s1], ph1)):
                    while True:
                        try:
                            wait(bar, par)
                            break
                        except Blocked:
                            yield "blocked"
                issue_stage(s, tile, chunks[j], buf, j == 0)
                issue_stage(s1, tile, chunks[j + 1], buf, False)
                s, ph = s1 + 1, ph1
                if s == S:
                    s, ph = 0, ph ^ 1
                j += 2
                yield "trip"
            while j < cnt:
                for bar, par in ((self.full[s], ph),):
                    while True:
                        try:
                            wait(bar, par)
                            break
                        except Blocked:
                            yield "blocked"
                issue_stage(s, tile, chunks[j], buf, j == 0)
                s += 1
                if s == S:
                    s, ph = 0, ph ^ 1
                j += 1
                yield "trip"
            self.pipe.append(("commit", self.tmem_full[buf], []))
            tl += 1

    def epi(self):
        tl = 0
        while True:
            buf = tl & 1
            while True:
                try:
                    wait(self.tmem_full[buf], (tl >> 1) & 1)
                    break
                except Blocked:
                    yield "blocked"
            tile = self.s_epi[buf]
            if tile < 0:
                return
            yield "ld"
            assert self.acc[buf] == list(self.sim.tiles[tile]), ("epilogue saw a wrong accumulation", tile, self.acc[buf])
            assert tile not in self.sim.done_tiles, ("tile processed twice", tile)
            self.sim.done_tiles.add(tile)
            self.acc_busy[buf] = False
            self.tmem_empty[buf].arrive()
            tl += 1

    # ---- asynchronous hardware ------------------------------------------------------------------------------------
    def step_pipe(self):
        """Retire the oldest item of the tensor pipe (in order)."""
        item = self.pipe.pop(0)
        if item[0] == "mma":
            _, sa, tag, buf, first = item
            assert self.a_slot[sa] == tag, ("MMA read a wrong A slot", sa, self.a_slot[sa], tag)
            assert self.b_slot[sa] == tag, ("MMA read a wrong weight slot", sa, self.b_slot[sa], tag)
            if first:
                self.acc[buf] = []
            self.acc[buf].append(tag[1])
            self.a_slot[sa] = ("read", tag)                   # consumed, but not yet released by a commit
            self.b_slot[sa] = ("read", tag)
        else:
            _, bar, slots = item
            for sl in slots:                                  # a commit releases every stage issued before it
                assert self.a_slot[sl] is None or self.a_slot[sl][0] == "read", ("commit before its MMA", sl, self.a_slot[sl])
                self.a_slot[sl] = None
                self.b_slot[sl] = None
            bar.arrive()


class Sim:
    def __init__(self, p, tiles, grid, seed):
        self.p, self.tiles, self.num_tiles, self.grid = p, tiles, len(tiles), grid
        self.P = min(grid, self.num_tiles)
        self.counter, self.fetches, self.done_tiles = 0, 0, set()
        self.rng = random.Random(seed)
        self.ctas = [CTA(c, self) for c in range(self.P)]      # CTAs without a first tile leave at once

    def run(self, max_steps=2_000_000):
        rng = self.rng
        live = [(c, name) for c in self.ctas for name in c.agents]
        idle_rounds = 0
        for _ in range(max_steps):
            if not live:
                break
            progressed = False
            # hardware first, sometimes: retire pipe items / land TMA copies in random order
            for c in self.ctas:
                if c.pipe and rng.random() < 0.5:
                    c.step_pipe()
                    progressed = True
                if c.async_ev and rng.random() < 0.5:
                    c.async_ev.pop(rng.randrange(len(c.async_ev)))()
                    progressed = True
            c, name = live[rng.randrange(len(live))]
            try:
                r = next(c.agents[name])
                if r != "blocked":
                    progressed = True
            except StopIteration:
                live.remove((c, name))
                progressed = True
            if progressed:
                idle_rounds = 0
            else:
                idle_rounds += 1
                if idle_rounds > 50 * (len(live) + 1):
                    # nothing can move unless hardware still has work: flush it, else it is a deadlock
                    hw = any(c.pipe or c.async_ev for c in self.ctas)
                    if not hw and self._all_blocked(live):
                        raise AssertionError("deadlock: %s" % sorted(n for _, n in live))
                    for c in self.ctas:
                        while c.pipe:
                            c.step_pipe()
                        while c.async_ev:
                            c.async_ev.pop()()
                    idle_rounds = 0
        else:
            raise AssertionError("model did not terminate")
        for c in self.ctas:
            while c.pipe:
                c.step_pipe()
        assert self.done_tiles == set(range(self.num_tiles)), sorted(set(range(self.num_tiles)) - self.done_tiles)
        if self.p["DYN"]:
            assert self.counter == 0 and self.fetches == self.num_tiles, (self.counter, self.fetches, self.num_tiles)

    def _all_blocked(self, live):
        for c, name in live:
            try:
                if next(c.agents[name]) != "blocked":
                    return False
            except StopIteration:
                return False
        return True


def _tiles(rng, n, T, style):
    out = []
    for t in range(n):
        if style == "ones":
            k = 1
        elif style == "sparse":
            k = rng.choice([1, 1, 2, 3])
        elif style == "full":
            k = T
        else:
            k = rng.randint(1, T)
        out.append(sorted(rng.sample(range(T), k)))
    return out


CONFIGS = [
    dict(STAGES=4, G=4, CG=1, DEPTH=2, DYN=1),     # shipped: N = 64, 16 producer warps, fp32 rows
    dict(STAGES=6, G=4, CG=1, DEPTH=2, DYN=1),     # shipped: N = 32, fp32 rows
    dict(STAGES=8, G=4, CG=1, DEPTH=2, DYN=1),     # shipped: N = 64, split rows (rings twice as deep)
    dict(STAGES=12, G=4, CG=1, DEPTH=2, DYN=1),    # shipped: N = 32, split rows
    dict(STAGES=8, G=2, CG=1, DEPTH=2, DYN=1),     # shipped: N = 128, split rows (8 producer warps)
    dict(STAGES=4, G=2, CG=1, DEPTH=4, DYN=1),     # 8 producer warps (N = 128 uses DEPTH 2)
    dict(STAGES=4, G=4, CG=1, DEPTH=2, DYN=0),     # static tiles
    dict(STAGES=4, G=4, CG=2, DEPTH=2, DYN=1),     # experimental commit groups
    dict(STAGES=6, G=4, CG=2, DEPTH=2, DYN=1),
    dict(STAGES=6, G=4, CG=3, DEPTH=2, DYN=1),
    dict(STAGES=4, G=2, CG=2, DEPTH=2, DYN=0),
]


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: "S%(STAGES)d-G%(G)d-CG%(CG)d-D%(DEPTH)d-dyn%(DYN)d" % c)
@pytest.mark.parametrize("style", ["random", "ones", "sparse", "full"])
def test_protocol_is_deadlock_free_and_consistent(cfg, style):
    for seed in range(12):
        rng = random.Random(1000 * seed + 7)
        grid = rng.choice([1, 2, 3])
        n_tiles = rng.choice([1, 2, 3, 5, 9, 14])
        T = rng.choice([4, 14, 27])
        Sim(cfg, _tiles(rng, n_tiles, T, style), grid, seed).run()


def test_model_reproduces_the_blocking_iterator_deadlock():
    """Why the producers' iterator must not block on a later tile's list while gathers are in flight (the rule in
    conv_fwd_tc_kernel's `locate`): without it, runs of 1-chunk tiles deadlock — the bug a hardware run hit in round 1."""
    cfg = dict(STAGES=4, G=4, CG=1, DEPTH=2, DYN=1, ALWAYS_BLOCK=1)
    dead = 0
    for seed in range(20):
        rng = random.Random(seed)
        try:
            Sim(cfg, _tiles(rng, 20, 14, rng.choice(["ones", "sparse"])), 2, seed).run(max_steps=300000)
        except AssertionError as e:
            dead += "deadlock" in str(e)
    assert dead >= 15, dead


def test_model_detects_a_broken_protocol():
    """Sanity of the checker itself: a commit that overtakes the last MMA stage of its group must trip the model."""
    cfg = dict(STAGES=4, G=4, CG=2, DEPTH=2, DYN=1)
    bad = 0
    for seed in range(10):
        try:
            _run_with_early_commit(cfg, _tiles(random.Random(seed), 9, 14, "random"), 2, seed)
        except AssertionError:
            bad += 1
    assert bad >= 8, bad


def _run_with_early_commit(cfg, tiles, grid, seed):
    sim = Sim(cfg, tiles, grid, seed)
    for c in sim.ctas:
        real = c.pipe

        class EarlyCommit(list):
            def append(self, item, _c=c):
                if item[0] == "commit" and item[2]:
                    # release the group when only its first stage has been issued
                    item = ("commit", item[1], item[2])
                    list.insert(self, max(len(self) - 1, 0), item)
                else:
                    list.append(self, item)
        c.pipe = EarlyCommit(real)
    sim.run()
