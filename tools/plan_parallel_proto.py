#!/usr/bin/env python
"""Data-parallel formulation of the partition planner (prototype for the device-side planner of DESIGN.md section 7, item 2).

The host planner (dskgpu.cu::plan_partitions_host) packs consecutive minimizer bins greedily: a partition is closed before
bin b when it already holds k-mers and bin b would take it beyond T.  That loop is sequential and memory-bound on the host
(2 x 2^level x 16 B per job).  The same plan, bit for bit, comes out of three data-parallel steps:

  1. prefix sums of the k-mers per bin (cum / excl);
  2. next[s] = where the partition that starts at bin s ends: one binary search per bin,
       b0 = first b with cum[b] > excl[s] + T ;  next[s] = b0 if the partition is non-empty before b0 else b0 + 1;
  3. the partition starts are the orbit of bin 0 under `next`: pointer doubling (log2 steps), marking from the highest
     power down;  partition id = inclusive scan of the start flags - 1.

Heavy partitions (beyond the shared-memory path) are then renumbered to the end, heaviest first (a stable sort of a few
hundred entries).  tests/test_host_logic.py::test_parallel_planner_prototype_equals_the_host_planner checks this prototype
against the C++ planner through dskgpu_selftest_plan.  numpy stands in for the device primitives (scan, vectorised binary
search, gather)."""
import numpy as np


def plan_parallel(km, T, lim=None, world=1):
    """km: uint64[NB] whole-job k-mers per bin.  Returns (bin2part uint32[NB], part_kmers uint64[P padded to world])."""
    km = np.asarray(km, dtype=np.uint64)
    NB = km.size
    cum = np.cumsum(km, dtype=np.uint64)
    excl = cum - km
    # step 2: end of the partition that would start at every bin
    b0 = np.searchsorted(cum, excl + np.uint64(T), side="right")            # first b with cum[b] > excl[s] + T  (NB if none)
    b0c = np.minimum(b0, NB - 1)
    empty_before = (b0 < NB) & (excl[b0c] == excl)                           # nothing but empty bins in [s, b0)
    nxt = np.where(empty_before, b0 + 1, b0).astype(np.int64)
    nxt = np.minimum(nxt, NB)                                                # NB = terminal
    # step 3: orbit of 0 under next, by pointer doubling over the table extended with the terminal node
    jump = np.concatenate([nxt, [NB]])                                       # jump[NB] = NB
    levels = [jump]
    while (1 << len(levels)) < NB + 1:
        j = levels[-1]
        levels.append(j[j])
    marked = np.zeros(NB + 1, dtype=bool)
    marked[0] = True
    for j in reversed(levels):                                               # highest power first
        marked[j[np.nonzero(marked)[0]]] = True
    start = marked[:NB]
    b2p = (np.cumsum(start) - 1).astype(np.uint32)
    P = int(b2p[-1]) + 1
    pk = np.zeros(P, dtype=np.uint64)
    np.add.at(pk, b2p, km)
    # heavy partitions to the end, heaviest first (stable), every rank from the same whole-job sums
    if lim is not None:
        heavy = np.nonzero(pk > np.uint64(lim))[0]
        if 0 < heavy.size < P:
            order = heavy[np.argsort(-pk[heavy].astype(np.int64), kind="stable")]
            newid = np.full(P, -1, dtype=np.int64)
            newid[order] = P - heavy.size + np.arange(heavy.size)
            light = np.nonzero(newid < 0)[0]
            newid[light] = np.arange(light.size)
            pk2 = np.zeros(P, dtype=np.uint64)
            pk2[newid] = pk
            pk = pk2
            b2p = newid[b2p].astype(np.uint32)
    Ppad = (P + world - 1) // world * world
    pk = np.concatenate([pk, np.zeros(Ppad - P, dtype=np.uint64)])
    return b2p, pk


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    km = (rng.pareto(1.3, 1 << 16) * 3000).astype(np.uint64)
    b2p, pk = plan_parallel(km, 30000, 88000, 2)
    print("bins", km.size, "partitions", pk.size, "largest", int(pk.max()))
