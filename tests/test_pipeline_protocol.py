"""CPU check of the forward kernel's mbarrier protocol (tools/protocol_sim.py: a discrete-event model of the TMA producer /
MMA issuer / epilogue warpgroups of vqb_fwd_tc_kernel) -- no deadlock, no missed barrier phase, no ring slot overwritten
under a reader, every tile processed exactly once, over many random interleavings.  Covers the shipped configurations
and the variants written after the last GPU visit of round 1 (not yet run on hardware)."""
import os
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import protocol_sim as PS  # noqa: E402

# kernel template arguments -> model parameters (PIECES per chunk: hi/lo per K-block + the bias block)
SHIPPED = {
    "p_code D=64 (two warpgroups, resident codebook)": dict(XS=2, BS=5, PIECES=5, chunks=1, NWG=2, RESIDENT=True, PASSES=3, PCODE=True),
    "p_code D=32": dict(XS=2, BS=3, PIECES=3, chunks=1, NWG=2, RESIDENT=True, PASSES=3, PCODE=True),
    "resident search K<=128 D=64 (one x slot)": dict(XS=1, BS=5, PIECES=5, chunks=1, RESIDENT=True, PASSES=3),
    "resident search K<=128 D=32": dict(XS=2, BS=3, PIECES=3, chunks=1, RESIDENT=True, PASSES=3),
    "streamed 3xTF32 D=64 K=1024": dict(XS=2, BS=4, PIECES=5, chunks=8, PASSES=3),
    "streamed 3xTF32 D=32 K=1024": dict(XS=2, BS=8, PIECES=3, chunks=8, PASSES=3),
    "streamed 3xTF32 D=128 (one x slot)": dict(XS=1, BS=4, PIECES=9, chunks=4, PASSES=3),
    "streamed 3xTF32 D=64 K=256 pipelined x_lo": dict(XS=2, BS=4, PIECES=5, chunks=2, PASSES=3, PIPE=True),
    "streamed 3xTF32 D=32 K=200 pipelined x_lo": dict(XS=2, BS=8, PIECES=3, chunks=2, PASSES=3, PIPE=True),
    "streamed 1xTF32 D=64 K=4096": dict(XS=2, BS=8, PIECES=3, chunks=32, PASSES=1),
    "streamed 1xTF32 D=256 K=8192 (NOAUG)": dict(XS=1, BS=5, PIECES=8, chunks=64, PASSES=1),
}
EXPERIMENTAL = {
    "column split D=64": dict(XS=2, BS=7, PIECES=3, chunks=16, PASSES=1, CS=2),
    "column split D=128 (one x slot)": dict(XS=1, BS=7, PIECES=5, chunks=8, PASSES=1, CS=2),
    "multicast pair D=256": dict(XS=1, BS=5, PIECES=8, chunks=8, PASSES=1, MC=2),
    "multicast pair D=128": dict(XS=1, BS=8, PIECES=5, chunks=8, PASSES=1, MC=2),
    "three-slot p_code D=64": dict(XS=3, BS=4, PIECES=4, chunks=1, NWG=2, RESIDENT=True, PASSES=3, PCODE=True),
    "three-slot p_code D=32": dict(XS=3, BS=2, PIECES=2, chunks=1, NWG=2, RESIDENT=True, PASSES=3, PCODE=True),
}


@pytest.mark.parametrize("name", list(SHIPPED) + list(EXPERIMENTAL))
@pytest.mark.parametrize("tiles,grid", [(1, 2), (2, 2), (5, 2), (7, 2), (12, 4), (13, 4)])
def test_protocol_completes(name, tiles, grid):
    cfg = dict((SHIPPED.get(name) or EXPERIMENTAL[name]), tiles=tiles, grid=grid)
    if cfg.get("MC") != 2 and tiles < grid:
        cfg["grid"] = tiles                                    # the launcher never starts more CTAs than tiles
    assert PS.check(cfg, seeds=25) is None


def test_pipelined_x_lo_beyond_two_chunks_deadlocks_as_on_the_gpu():
    """What the B200 showed when the pipelined x_lo was forced on at K = 1024 (8 chunks per tile): the epilogue waits for
    x(t+1) at the top of tile t, the producer issues x(t+1) only after the last codebook piece of tile t, the ring frees
    slots only as the MMA consumes them, and the MMA stops two chunks ahead of the epilogue.  The kernel guards the
    variant with num_chunks <= 2; the model shows the same cycle, and that two chunks are safe."""
    bad = dict(XS=2, BS=4, PIECES=5, chunks=8, PASSES=3, PIPE=True, tiles=6, grid=2)
    assert "blocked" in (PS.check(bad, seeds=5) or "")
    ok = dict(bad, chunks=2)
    assert PS.check(ok, seeds=25) is None
    # three chunks already close the cycle whenever the ring is shorter than the pieces of the third chunk onwards
    assert "blocked" in (PS.check(dict(bad, chunks=3), seeds=10) or "")


def test_model_notices_a_wrong_barrier_count():
    """Sanity of the model itself: the multicast pair with b_empty left at one arrival (the non-cluster count) lets a slot
    be refilled while the peer still reads it, or loses a phase -- the model must not report that as fine."""
    cfg = dict(XS=1, BS=5, PIECES=8, chunks=8, PASSES=1, MC=2, tiles=6, grid=2)
    real_init = PS.Cta.__init__

    def broken_init(self, c, rank=0):
        real_init(self, c, rank)
        self.b_empty = [PS.MBar("b_empty%d[%d]" % (rank, i), 1) for i in range(c["BS"])]

    PS.Cta.__init__ = broken_init
    try:
        failed = False
        try:
            failed = PS.check(cfg, seeds=25) is not None
        except AssertionError:
            failed = True
        assert failed
    finally:
        PS.Cta.__init__ = real_init


def test_model_notices_an_x_lo_slot_rewritten_under_the_mma():
    """Two warpgroups share ONE x_lo tile; without the xlo_free hand-back the second warpgroup rewrites it while the third
    MMA pass of the previous tile may still read it."""
    cfg = dict(SHIPPED["p_code D=64 (two warpgroups, resident codebook)"], tiles=8, grid=2, BUG_skip_xlo_free=True)
    failed = False
    try:
        failed = PS.check(cfg, seeds=40) is not None
    except AssertionError as e:
        failed = "x_lo slot" in str(e)
    assert failed


@pytest.mark.parametrize("chunks,BS,PIECES", [(8, 4, 5), (3, 4, 5), (8, 8, 3), (32, 4, 5), (2, 4, 5)])
def test_design_study_late_prepared_x_lo_is_deadlock_free_for_any_chunk_count(chunks, BS, PIECES):
    """Design study for round 2 (not in the kernel): producing x_lo of tile t+1 AFTER the chunk loop of tile t -- i.e.
    under the re-rank, gather and store of tile t only -- has no cycle whatever the number of chunks, because by then
    the MMA has retired every piece of tile t and the producer has moved on to x(t+1)."""
    cfg = dict(XS=2, BS=BS, PIECES=PIECES, chunks=chunks, PASSES=3, PIPE=True, PIPE_LATE=True, tiles=9, grid=2)
    assert PS.check(cfg, seeds=25) is None
