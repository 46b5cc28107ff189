"""Multi-GPU orchestration of the counting path: one process per GPU, torch.distributed for the plumbing.

Every rank parses its own slice of the reads; super-k-mer records are routed to the rank owning their partition
(owner(p) = p % world_size) by the partition-scatter kernel itself, which stores straight into the owners' HBM
through CUDA-IPC peer pointers (NVLink P2P) -- the role the reference gives to its temp files
(SuperKmerBinFiles, G/src/gatb/tools/storage/impl/Storage.cpp:310-589).  Only tiny metadata crosses
torch.distributed: four job totals and the minimizer-bin histogram (all-reduce, 1-16 MB), the per-partition count matrix (all-gather) and
the IPC handles.
"""
import ctypes as C

import numpy as np

from . import _lib


def exchange_layout(world_size, all_counts, sender):
    """offsets[p] of `sender`'s records inside owner(p)'s receive buffer + records each rank receives (host only)."""
    all_counts = np.ascontiguousarray(all_counts, dtype=np.uint64)
    P = all_counts.shape[1] // 2
    off = np.zeros(P, dtype=np.uint64)
    recv = np.zeros(world_size, dtype=np.uint64)
    rc = _lib.lib().dskgpu_xchg_layout(world_size, P, all_counts.ctypes.data, sender, off.ctypes.data, recv.ctypes.data)
    if rc != 0:
        raise RuntimeError("dskgpu_xchg_layout failed: %d" % rc)
    return off, recv


def all_gather_counts(dist, counts, device=None):
    """all-gather a uint64 numpy vector over torch.distributed (NCCL on `device`, gloo on CPU)."""
    import torch
    t = torch.from_numpy(counts.astype(np.int64))
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return np.stack([o.cpu().numpy().astype(np.uint64) for o in out])


def _grow_plan(need, announced):
    """capacities (records) every rank's receive buffer should have: a rank whose need exceeds what it announced grows to
    125 % of the need.  Pure function of data every rank holds, so all ranks take the same decision."""
    return [int(a) if int(n) <= int(a) else int(n) + int(n) // 4 + 1024 for n, a in zip(need, announced)]


def distributed_finish(eng, dist, device):
    """Runs the exchange + local counting on every rank (call after the pushes).

    Protocol (include/dskgpu.h "exchange v2"): all-reduce of four job totals, all-reduce of the bin histogram and
    all-gather of the per-partition record counts on DEVICE buffers (NCCL), CUDA-IPC handles only when a receive buffer
    had to grow, then the records cross NVLink as whole partition segments and every rank counts what it owns."""
    import os
    import sys
    import time
    import torch
    tr = os.environ.get("DSKGPU_TRACE_XCHG") and (dist.get_rank() == 0 or os.environ.get("DSKGPU_TRACE_XCHG") == "all")
    marks = [("start", time.perf_counter())]

    def mark(name):
        if tr:
            marks.append((name, time.perf_counter()))

    W, rank = dist.get_world_size(), dist.get_rank()
    cur = torch.cuda.current_stream(device)
    # cfg.stream == NULL (also what torch reports for the legacy default stream) means a library-OWNED non-blocking
    # stream: it is never "the same stream" as torch's, so the two sides must be ordered by explicit syncs.
    same_stream = bool(eng.stream) and int(eng.stream) == int(cur.cuda_stream)

    def to_torch():          # work queued by the library must be visible to torch's stream
        if not same_stream:
            eng.xchg_sync()

    def to_lib():            # ... and the other way round
        if not same_stream:
            cur.synchronize()

    g4 = torch.from_numpy(eng.xchg_prepare().astype(np.int64)).to(device)     # k-mers, records, density sample (k-mers, distinct)
    mark("prepare (push kernels done)")
    dist.all_reduce(g4)                                                       # every rank picks the same bin level / partition size
    level = eng.xchg_set_global(g4.cpu().numpy().astype(np.uint64))
    mark("allreduce totals")
    G = torch.empty(2 << level, dtype=torch.int64, device=device)             # (records, k-mers) per minimizer bin
    eng.xchg2_hist(G.data_ptr())
    to_torch()
    dist.all_reduce(G)                                                         # every rank plans the same partitions
    to_lib()
    counts, need = eng.xchg2_plan(G.data_ptr())
    mark("histogram allreduce + plan")
    mine = torch.from_numpy(counts.view(np.int64)).to(device)
    M = torch.empty(W * mine.numel(), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(M, mine)                                       # [W][P] on the device; the host never reads it
    mark("allgather counts (async)")
    # receive buffers: handles travel only when somebody has to grow (every rank sees the same `need`)
    st = getattr(eng, "_xchg_state", None)
    if st is None:
        st = eng._xchg_state = {"announced": [0] * W, "ptrs": [0] * W, "handles": [None] * W}
    want = _grow_plan(need, st["announced"])
    if want != st["announced"]:
        eng.xchg2_ensure_recv(want[rank])
        handle = np.frombuffer(eng.xchg_ipc_handle(), dtype=np.uint8).copy()
        hs = all_gather_counts(dist, np.frombuffer(handle.tobytes(), dtype=np.uint64), device)      # 8 x u64 per rank
        for r in range(W):
            hb = hs[r].tobytes()
            if r != rank and st["handles"][r] != hb:
                st["handles"][r] = hb
                st["ptrs"][r] = eng.xchg_open_peer(hb)
        st["announced"] = want
    else:
        eng.xchg2_ensure_recv(want[rank])
    eng.xchg_set_peers(st["ptrs"])
    mark("receive buffers / handles")
    to_lib()
    eng.xchg2_scatter(M.data_ptr())
    eng.xchg_sync()
    mark("scatter + segment copy")
    dist.barrier()                       # every rank's records have landed before anyone counts
    mark("barrier")
    eng.finish()
    mark("finish (count + order)")
    if tr:
        for (_, a), (name, b) in zip(marks, marks[1:]):
            sys.stderr.write("[xchg r%d] %-32s +%8.3f ms\n" % (dist.get_rank(), name, 1e3 * (b - a)))
    return M


def in_process_finish(engines):
    """Same protocol for several contexts living in ONE process (tests: N 'ranks' on one GPU, no IPC needed)."""
    W = len(engines)
    g4 = np.sum([e.xchg_prepare() for e in engines], axis=0, dtype=np.uint64)
    for e in engines:
        e.xchg_set_global(g4)
    gh = np.sum([e.xchg_bin_hist() for e in engines], axis=0, dtype=np.uint64)
    allc = np.stack([e.xchg_part_counts(gh) for e in engines])
    for e in engines:
        e.xchg_plan(allc)
    ptrs = [e.xchg_recv_buffer()[0] for e in engines]
    for e in engines:
        e.xchg_set_peers(ptrs)
        e.xchg_scatter()
    for e in engines:
        e.xchg_sync()
    for e in engines:
        e.finish()
    return allc


def in_process_finish_v2(engines, device="cuda"):
    """exchange v2 for several contexts living in ONE process (tests): torch sums / stacks stand in for NCCL."""
    import torch
    g4 = np.sum([e.xchg_prepare() for e in engines], axis=0, dtype=np.uint64)
    levels = [e.xchg_set_global(g4) for e in engines]
    assert len(set(levels)) == 1
    hs = [torch.empty(2 << levels[0], dtype=torch.int64, device=device) for _ in engines]
    for e, h in zip(engines, hs):
        e.xchg2_hist(h.data_ptr())
        e.xchg_sync()
    G = torch.stack(hs).sum(0)
    torch.cuda.synchronize()
    plans = [e.xchg2_plan(G.data_ptr()) for e in engines]
    M = torch.from_numpy(np.stack([c for c, _ in plans]).view(np.int64)).to(device).contiguous()
    torch.cuda.synchronize()
    for e, (_, need) in zip(engines, plans):
        assert (need == plans[0][1]).all()
        e.xchg2_ensure_recv(int(need[e.cfg.rank]))
    ptrs = [e.xchg_recv_buffer()[0] for e in engines]
    for e in engines:
        e.xchg_set_peers(ptrs)
        e.xchg2_scatter(M.data_ptr())
    for e in engines:
        e.xchg_sync()
    for e in engines:
        e.finish()
    return M.cpu().numpy()
