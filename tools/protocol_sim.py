"""Discrete-event model of the mbarrier protocol of vqb_fwd_tc_kernel (semi-tts_b200/csrc/vqb_fwd_tc.cu) -- a CPU check for
deadlocks and missed barrier phases in the warp-specialised pipeline (TMA producer / MMA issuer / epilogue warpgroups).

It mirrors the CONTROL FLOW of the kernel, role by role and wait by wait (the comments name the statements), not its
arithmetic: tiles, x slots, the codebook ring, the two TMEM accumulator buffers, the x_lo hand-off of the 3xTF32 passes,
and the variants added late in round 1 -- the pipelined x_lo (PIPE), the column-split second epilogue warpgroup (CS = 2),
the cluster-of-two multicast codebook stream (MC = 2) and the three-slot p_code forward (XS = 3, NOAUG).  Asynchronous
completions (TMA bytes, tcgen05.commit arrivals) land after random delays and the runnable role is picked at random, so a
few hundred seeds walk many interleavings.  `mbarrier.try_wait.parity P` is modelled as "the number of completed phases
is odd when P == 0 / even when P == 1" -- a waiter that is lapped by two phases blocks forever, as on the hardware.

Used by tests/test_pipeline_protocol.py.  It reproduces what the B200 showed: the pipelined x_lo deadlocks with more than
two chunks per tile (the kernel therefore guards it with num_chunks <= 2).
"""
import random


class Deadlock(Exception):
    pass


class MBar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.tx, self.phases = name, count, count, 0, 0

    def _check(self):
        if self.pending == 0 and self.tx == 0:
            self.phases += 1
            self.pending = self.count
        assert self.pending >= 0, "barrier %s over-arrived" % self.name

    def arrive(self, n=1, tx=0):
        self.pending -= n
        self.tx += tx
        self._check()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        self._check()

    def passed(self, parity):                      # mbarrier.try_wait.parity
        return (self.phases & 1) != parity


class NamedBar:                                    # bar.sync id, n  (counted in warpgroups here)
    def __init__(self, n):
        self.n, self.arrived, self.gen = n, 0, 0


class Cta:
    def __init__(self, cfg, rank=0):
        c = cfg
        mk = lambda nm, n, cnt: [MBar("%s%d[%d]" % (nm, rank, i), cnt) for i in range(n)]
        self.x_full, self.x_empty, self.xlo_full = mk("x_full", c["XS"], 1), mk("x_empty", c["XS"], 4), mk("xlo_full", c["XS"], 4)
        self.b_full, self.b_empty = mk("b_full", c["BS"], 1), mk("b_empty", c["BS"], c["MC"])
        self.t_full, self.t_empty = mk("t_full", 2, 1), mk("t_empty", 2, 4 * c["CS"])
        self.xlo_free = MBar("xlo_free%d" % rank, 1)
        self.nb3, self.nb4 = NamedBar(2), NamedBar(2)
        self.rank = rank
        self.slot_busy = [False] * c["BS"]         # ring slot holds a piece the MMA of THIS CTA has not retired yet
        self.xlo_reading = [False] * c["XS"]       # the MMA's third pass may still read this x_lo slot

    def write_xlo(self, xls):                      # the epilogue warps overwrite an x_lo slot, then arrive on xlo_full
        assert not self.xlo_reading[xls], "x_lo slot %d of CTA %d rewritten under the MMA's third pass" % (xls, self.rank)
        self.xlo_full[xls].arrive(4)

    def land(self, bs):                            # TMA bytes of a codebook piece arrive in ring slot bs
        assert not self.slot_busy[bs], "ring slot %d of CTA %d overwritten while its MMA may still read it" % (bs, self.rank)
        self.slot_busy[bs] = True
        self.b_full[bs].complete_tx(1)


class Sim:
    def __init__(self, cfg, seed=0):
        d = dict(XS=2, BS=4, PIECES=5, chunks=2, tiles=6, grid=2, NWG=1, RESIDENT=False, PASSES=3, PCODE=False,
                 PIPE=False, CS=1, MC=1)
        d.update(cfg)
        self.c = d
        self.rng = random.Random(seed)
        self.async_q = []                          # (due, fn)
        self.now = 0
        self.tasks = []
        n_cta = d["grid"]
        self.ctas = [Cta(d, r) for r in range(n_cta)]
        self.done_tiles = {r: [] for r in range(n_cta)}
        for r in range(n_cta):
            self.tasks.append(("producer%d" % r, self.producer(r)))
            self.tasks.append(("mma%d" % r, self.mma(r)))
            for g in range(d["NWG"] * d["CS"]):
                self.tasks.append(("epi%d.%d" % (r, g), self.epilogue(r, g)))

    # ---- helpers -------------------------------------------------------------------------------------------------
    def my_tiles(self, rank):
        c = self.c
        if c["MC"] == 2:                           # tile_end: ceil(num_tiles / grid) iterations for every CTA
            n_iter = -(-c["tiles"] // c["grid"])
            return [rank + i * c["grid"] for i in range(n_iter)]
        return list(range(rank, c["tiles"], c["grid"]))

    def later(self, fn, lo=1, hi=12):
        self.async_q.append((self.now + self.rng.randint(lo, hi), fn))

    def peer(self, rank):
        return self.ctas[rank ^ 1]

    # ---- roles ---------------------------------------------------------------------------------------------------
    def producer(self, rank):
        c, k = self.c, self.ctas[rank]
        x_it = b_it = 0
        resident_loaded = False
        for _tile in self.my_tiles(rank):
            xs, xph = x_it % c["XS"], (x_it // c["XS"]) & 1
            yield ("wait", k.x_empty[xs], xph ^ 1)                       # mbar_wait(&x_empty[xs], xph ^ 1)
            k.x_full[xs].arrive(1, tx=1)                                 # mbar_arrive_expect_tx
            self.later(lambda b=k.x_full[xs]: b.complete_tx(1))          # tma_load_2d (all K-blocks)
            x_it += 1
            if c["RESIDENT"] and resident_loaded:
                continue
            resident_loaded = True
            for _chunk in range(c["chunks"]):
                for _j in range(c["PIECES"]):
                    bs, bph = b_it % c["BS"], (b_it // c["BS"]) & 1
                    if not c["RESIDENT"]:
                        yield ("wait", k.b_empty[bs], bph ^ 1)
                    k.b_full[bs].arrive(1, tx=1)
                    if c["MC"] == 2:
                        if (b_it & 1) == (rank & 1):                     # this CTA multicasts the piece into both ring slots
                            self.later(lambda a=k, b=self.peer(rank), i=bs: (a.land(i), b.land(i)))
                    else:
                        self.later(lambda a=k, i=bs: a.land(i))
                    b_it += 1
                    yield ("step",)

    def mma(self, rank):
        c, k = self.c, self.ctas[rank]
        x_it = b_it = c_it = 0
        resident_ready = False
        commits = []                                                     # tcgen05.commit arrivals retire in issue order

        def commit(*bars, frees=None, xlo_done=None):
            due = max([self.now] + [d for d, _ in commits]) + self.rng.randint(1, 6)

            def fn(bs=bars, fr=frees, xd=xlo_done):
                if fr is not None and not c["RESIDENT"]:
                    k.slot_busy[fr] = False
                if xd is not None:
                    k.xlo_reading[xd] = False
                for b in bs:
                    b.arrive(1)
            commits.append((due, fn))
            self.async_q.append((due, fn))

        for _tile in self.my_tiles(rank):
            xs, xph = x_it % c["XS"], (x_it // c["XS"]) & 1
            xls, xlph = (0, x_it & 1) if c["NWG"] == 2 else (xs, xph)
            yield ("wait", k.x_full[xs], xph)
            if c["PASSES"] == 3 and not c["RESIDENT"]:
                yield ("wait", k.xlo_full[xls], xlph)
                k.xlo_reading[xls] = True
            for chunk_i in range(c["chunks"]):
                buf, tph = c_it & 1, (c_it >> 1) & 1
                yield ("wait", k.t_empty[buf], tph ^ 1)
                if c["RESIDENT"]:
                    if not resident_ready:
                        for j in range(c["PIECES"]):
                            yield ("wait", k.b_full[j], 0)
                        resident_ready = True
                    if c["PASSES"] == 3:
                        yield ("wait", k.xlo_full[xls], xlph)
                        k.xlo_reading[xls] = True
                        if c["NWG"] == 2:
                            commit(k.xlo_free, xlo_done=xls)
                else:
                    for _j in range(c["PIECES"]):
                        bs, bph = b_it % c["BS"], (b_it // c["BS"]) & 1
                        yield ("wait", k.b_full[bs], bph)
                        if c["MC"] == 2:
                            commit(k.b_empty[bs], self.peer(rank).b_empty[bs], frees=bs)   # umma_commit_mc, mask 0b11
                        else:
                            commit(k.b_empty[bs], frees=bs)
                        b_it += 1
                last = chunk_i == c["chunks"] - 1 and c["PASSES"] == 3 and c["NWG"] != 2
                commit(k.t_full[buf], xlo_done=xls if last else None)
                c_it += 1
            x_it += 1

    def epilogue(self, rank, g):
        c, k = self.c, self.ctas[rank]
        wg, cg = (0, g) if c["CS"] == 2 else (g, 0)
        tiles = self.my_tiles(rank)[wg::c["NWG"]]
        x_it = c_it = wg
        pipe = c["PIPE"] and not c["PCODE"] and not c["RESIDENT"] and c["PASSES"] == 3 and c["XS"] == 2 and c["NWG"] == 1

        def prep(it):                                                    # prep_tile<KB, XS>
            xs, xph = it % c["XS"], (it // c["XS"]) & 1
            yield ("wait", k.x_full[xs], xph)
            k.write_xlo(xs)

        if pipe and tiles:
            yield from prep(x_it)
        for n, tile in enumerate(tiles):
            xs, xph = x_it % c["XS"], (x_it // c["XS"]) & 1
            xls, xlph = (0, x_it & 1) if c["NWG"] == 2 else (xs, xph)
            late = pipe and c.get("PIPE_LATE")                         # design study: prepare tile t+1 AFTER the chunk loop of tile t
            if pipe:
                if n + 1 < len(tiles) and not late:
                    yield from prep(x_it + 1)
            else:
                yield ("wait", k.x_full[xs], xph)
                if c["NWG"] == 2 and not c.get("BUG_skip_xlo_free"):         # (the BUG_ key exists for the model's own test)
                    yield ("wait", k.xlo_free, xlph ^ 1)
                if c["PASSES"] == 3:
                    k.write_xlo(xls)
            if c["PCODE"]:
                buf, tph = c_it & 1, (c_it >> 1) & 1
                yield ("wait", k.t_full[buf], tph)
                k.t_empty[buf].arrive(4)
                c_it += c["NWG"]
            else:
                for _chunk in range(c["chunks"]):
                    buf, tph = c_it & 1, (c_it >> 1) & 1
                    yield ("wait", k.t_full[buf], tph)
                    yield ("step",)
                    k.t_empty[buf].arrive(4)
                    c_it += 1
            if late and n + 1 < len(tiles):
                yield from prep(x_it + 1)
            if c["CS"] == 2:
                yield ("named", k.nb3)
                if cg == 1:
                    yield ("named", k.nb4)
                    x_it += c["NWG"]
                    continue
                yield ("step",)                                          # merge + re-rank over both lists
                yield ("named", k.nb4)
            yield ("step",)                                              # gather, stores
            self.done_tiles[rank].append(tile)
            k.x_empty[xs].arrive(4)
            x_it += c["NWG"]

    # ---- scheduler -----------------------------------------------------------------------------------------------
    def run(self, max_steps=2_000_000):
        live = [[name, gen, None] for name, gen in self.tasks]           # [name, generator, blocked-on]
        steps = 0
        while live:
            steps += 1
            if steps > max_steps:
                raise Deadlock("no progress after %d steps" % max_steps)
            self.now += 1
            due = [e for e in self.async_q if e[0] <= self.now]
            if due:
                self.async_q = [e for e in self.async_q if e[0] > self.now]
                for _, fn in sorted(due, key=lambda e: e[0]):
                    fn()
            runnable = []
            for t in live:
                b = t[2]
                if b is None:
                    runnable.append(t)
                elif b[0] == "wait" and b[1].passed(b[2]):
                    t[2] = None
                    runnable.append(t)
                elif b[0] == "named" and b[1].gen != b[2]:
                    t[2] = None
                    runnable.append(t)
            if not runnable:
                if self.async_q:
                    self.now = min(e[0] for e in self.async_q) - 1
                    continue
                raise Deadlock("blocked: " + ", ".join("%s on %s" % (t[0], t[2][1].name if t[2][0] == "wait" else "bar.sync")
                                                       for t in live))
            t = self.rng.choice(runnable)
            try:
                op = next(t[1])
            except StopIteration:
                live.remove(t)
                continue
            if op[0] == "wait":
                if not op[1].passed(op[2]):
                    t[2] = op
            elif op[0] == "named":
                nb = op[1]
                nb.arrived += 1
                if nb.arrived == nb.n:
                    nb.arrived = 0
                    nb.gen += 1
                else:
                    t[2] = ("named", nb, nb.gen)
        return self.done_tiles


def check(cfg, seeds=40):
    """Runs `seeds` random interleavings; returns None if all complete with every real tile processed once, else the
    first failure as a string."""
    for s in range(seeds):
        sim = Sim(cfg, seed=s)
        try:
            done = sim.run()
        except Deadlock as e:
            return "seed %d: %s" % (s, e)
        got = sorted(t for r in done.values() for t in r if t < sim.c["tiles"])
        if got != list(range(sim.c["tiles"])):
            return "seed %d: tiles processed %s" % (s, got)
    return None
