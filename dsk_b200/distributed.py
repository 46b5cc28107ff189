"""Multi-GPU orchestration of the counting path: one process per GPU, torch.distributed for the plumbing.

Every rank parses its own slice of the reads; super-k-mer records are routed to the rank owning their partition
(owner(p) = p % world_size) -- the role the reference gives to its temp files (SuperKmerBinFiles,
G/src/gatb/tools/storage/impl/Storage.cpp:310-589).  The partitions are planned ON THE DEVICE by every rank from the
all-reduced minimizer-bin histogram (dsk_b200/csrc/plan.cuh); the local records are scattered into owner-major partition
order, so what crosses NVLink is ONE contiguous chunk per (sender, receiver) pair, stored by the library's own copy kernel
through CUDA-IPC peer pointers (no NCCL on the data path).  torch.distributed only carries metadata, as DEVICE tensors on
the stream the kernels run on: four job totals and the bin histogram (all-reduce), the per-partition record counts
(all-to-all of rows), the chunk sizes (all-gather), the IPC handles when a receive buffer had to grow, and a one-element
all-reduce that orders "every chunk has landed" before the counting kernels without a host-side barrier.
"""
import ctypes as C

import numpy as np

from . import _lib


def exchange_layout(world_size, parts_per_rank, counts, owner):
    """host mirror of rank `owner`'s receive layout (dskgpu_xchg_layout): counts[s] = rank s's records per partition in q
    order ([W][W * PW]).  Returns (region_base[W + 1], seg_off[W][PW])."""
    counts = np.ascontiguousarray(counts, dtype=np.uint64)
    W, PW = world_size, parts_per_rank
    assert counts.shape == (W, W * PW)
    base = np.zeros(W + 1, dtype=np.uint64)
    seg = np.zeros((W, PW), dtype=np.uint64)
    rc = _lib.lib().dskgpu_xchg_layout(W, PW, counts.ctypes.data, owner, base.ctypes.data, seg.ctypes.data)
    if rc != 0:
        raise RuntimeError("dskgpu_xchg_layout failed: %d" % rc)
    return base, seg


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
    """Runs the exchange + local counting on every rank (call after the pushes).  Protocol: include/dskgpu.h "multi-GPU
    exchange".  Host syncs of the step: the totals (push kernels drained), the four summed totals, the plan header."""
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

    evs = []

    def ev(name):            # device-side timeline of the metadata phases (elapsed times read after finish(), which syncs)
        e = torch.cuda.Event(enable_timing=True)
        e.record(cur)
        evs.append((name, e))

    # A rank whose input was rejected (or that failed in any other way) must not leave the others waiting in a collective:
    # its failure travels as a fifth "total" of the first all-reduce, and every rank raises.
    failure = None
    try:
        l4 = eng.xchg_prepare()                                                # k-mers, records, density sample (k-mers, distinct)
        sk0 = eng.xchg_sketch()                                                # distinct k-mers of the UNION of the ranks' samples
    except Exception as ex:                                                    # noqa: BLE001 -- re-raised below, on every rank
        failure, l4, sk0 = ex, np.zeros(4, np.uint64), np.zeros(4096, np.uint32)
    g4 = torch.from_numpy(np.concatenate([l4.astype(np.int64), [1 if failure else 0]])).to(device)
    mark("prepare (push kernels done)")
    ev("start")
    sk = torch.from_numpy(sk0.astype(np.int32)).to(device)
    dist.all_reduce(g4)                                                       # every rank picks the same bin level / partition size
    dist.all_reduce(sk, op=dist.ReduceOp.MAX)
    g4h = g4.cpu().numpy()
    if g4h[4]:
        raise failure if failure is not None else RuntimeError("distributed_finish: %d other rank(s) failed before the exchange" % int(g4h[4]))
    eng.xchg_set_sketch(sk.cpu().numpy().astype(np.uint32))
    level = eng.xchg_set_global(g4h[:4].astype(np.uint64))
    mark("allreduce totals")
    ev("totals_allreduce")
    G = torch.empty(2 << level, dtype=torch.int64, device=device)             # (records, k-mers) per minimizer bin
    eng.xchg_hist(G.data_ptr())
    to_torch()
    dist.all_reduce(G)                                                         # every rank plans the same partitions
    ev("hist_allreduce")
    to_lib()
    try:
        P, PW, need = eng.xchg_plan(G.data_ptr())                              # device planner; the host reads one header
    except Exception as ex:                                                    # noqa: BLE001
        failure, P, PW, need = ex, 0, 0, np.zeros(W, np.uint64)
    okf = torch.tensor([1 if failure else 0], dtype=torch.int32, device=device)
    dist.all_reduce(okf)                                                       # (out of memory while planning, ...: all ranks leave together)
    if int(okf.cpu()[0]):
        raise failure if failure is not None else RuntimeError("distributed_finish: another rank failed while planning")
    mark("histogram allreduce + device plan")
    ev("plan")
    send = torch.empty(W * PW + W, dtype=torch.int64, device=device)
    eng.xchg_counts(send.data_ptr())
    to_torch()
    rows = torch.empty(W * PW, dtype=torch.int64, device=device)               # row s = rank s's records of MY partitions
    dist.all_to_all_single(rows, send[:W * PW])
    S = torch.empty(W * W, dtype=torch.int64, device=device)                   # S[s][o] = records rank s holds for rank o
    dist.all_gather_into_tensor(S, send[W * PW:])
    mark("count rows all-to-all (async)")
    ev("rows_alltoall")
    # receive buffers: handles travel only when somebody has to grow (every rank sees the same `need`)
    st = getattr(eng, "_xchg_state", None)
    if st is None:
        st = eng._xchg_state = {"announced": [0] * W, "ptrs": [0] * W, "handles": [None] * W}
    want = _grow_plan(need, st["announced"])
    if want != st["announced"]:
        eng.xchg_ensure_recv(want[rank])
        handle = np.frombuffer(eng.xchg_ipc_handle(), dtype=np.uint8).copy()
        hs = all_gather_counts(dist, np.frombuffer(handle.tobytes(), dtype=np.uint64), device)      # 8 x u64 per rank
        for r in range(W):
            hb = hs[r].tobytes()
            if r != rank and st["handles"][r] != hb:
                if st["ptrs"][r]:
                    eng.xchg_close_peer(st["ptrs"][r])                        # the peer re-allocated: drop the stale mapping
                st["handles"][r] = hb
                st["ptrs"][r] = eng.xchg_open_peer(hb)
        st["announced"] = want
    else:
        eng.xchg_ensure_recv(want[rank])
    eng.xchg_set_peers(st["ptrs"])
    mark("receive buffers / handles")
    to_lib()
    eng.xchg_scatter(rows.data_ptr(), S.data_ptr())
    to_torch()
    ev("scatter_send")
    flag = torch.zeros(1, dtype=torch.int32, device=device)
    dist.all_reduce(flag)                # stream-ordered: completes on a rank only after every rank's copy kernel has finished
    ev("ordering_allreduce")
    to_lib()
    mark("scatter + chunk copies + ordering all-reduce (queued)")
    eng._xchg_keep = (G, send, rows, S, flag)                                  # alive until the kernels that read them have run
    eng.finish()
    mark("finish (count + order)")
    if same_stream:
        eng.xchg_times = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(evs, evs[1:])}
    if tr:
        for (_, a), (name, b) in zip(marks, marks[1:]):
            sys.stderr.write("[xchg r%d] %-52s +%8.3f ms\n" % (dist.get_rank(), name, 1e3 * (b - a)))
    return P
