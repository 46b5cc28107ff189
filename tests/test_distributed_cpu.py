"""N>1 host logic on CPU: world_size-2 gloo processes agree on the plan and on the sender-major receive layout the exchange
kernels use (dskgpu_selftest_plan = the sequential mirror of the device planner, dskgpu_xchg_layout = the mirror of
k_xchg_bases; include/dskgpu.h "multi-GPU exchange")."""
import os
import subprocess
import sys

import numpy as np

from dsk_b200.distributed import exchange_layout, _grow_plan

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from dsk_b200 import _lib
from dsk_b200.distributed import exchange_layout, all_gather_counts, _grow_plan
dist.init_process_group("gloo")
rank, W = dist.get_rank(), dist.get_world_size()
L = _lib.lib()
level = 16; nb = 1 << level
rng = np.random.default_rng(100 + rank)                      # every rank holds different records of every bin
lkm = (rng.pareto(1.3, nb) * 1500).astype(np.uint64)
lkm[rng.integers(0, nb, 3)] += np.uint64(1_500_000)          # hot minimizers: heavy partitions
lrec = lkm // np.uint64(11) + (lkm > 0).astype(np.uint64)
lh = np.concatenate([lrec, lkm])
g = torch.from_numpy(lh.astype(np.int64)); dist.all_reduce(g)    # the bin histogram all-reduce of the protocol
gh = g.numpy().astype(np.uint64)
b2p = np.zeros(nb, np.uint32); pk = np.zeros(nb + W, np.uint64); pl = np.zeros(nb + W, np.uint64)
P = L.dskgpu_selftest_plan(level, gh.ctypes.data, lh.ctypes.data, W, 1, 15360, 0.26, 0, 0, b2p.ctypes.data, pk.ctypes.data, pl.ctypes.data, pk.size)
assert P > 0
PW = (P + W - 1) // W
# this rank's records per partition in q order (q = (p % W) * PW + p // W): what dskgpu_xchg_counts hands to the all-to-all
cq = np.zeros(W * PW, np.uint64)
p = np.arange(P)
cq[(p % W) * PW + p // W] = pl[:P]
allc = all_gather_counts(dist, cq)                          # [W][W * PW]; the real protocol moves only row chunks (all-to-all)
plans = all_gather_counts(dist, np.concatenate([[np.uint64(P)], pk[:P]]))
assert (plans == plans[0]).all()                            # every rank derived the same plan from the same histogram
# every rank checks every owner's layout: the segments of the senders != owner tile the receive buffer exactly once, in
# sender-major order, and the owner's own chunk is addressed from 0
need = []
for o in range(W):
    base, seg = exchange_layout(W, PW, allc, o)
    recv = int(base[W])
    used = np.zeros(recv, np.int32)
    for s in range(W):
        cnt = allc[s][o * PW:(o + 1) * PW]
        if s == o:
            assert seg[s][0] == 0 and (np.diff(seg[s].astype(np.int64)) == cnt[:-1].astype(np.int64)).all()
            continue
        assert int(seg[s][0]) == int(base[s])               # region of sender s starts with its first owned job
        for j in range(PW):
            a = int(seg[s][j]); used[a:a + int(cnt[j])] += 1
        assert int(seg[s][PW - 1]) + int(cnt[PW - 1]) == int(base[s + 1])
    assert (used == 1).all(), o
    assert recv == sum(int(allc[s][o * PW:(o + 1) * PW].sum()) for s in range(W) if s != o)
    need.append(recv)
# receive-buffer growth is decided from numbers every rank holds
want = _grow_plan(need, [0] * W)
assert all(w >= n for w, n in zip(want, need)) and _grow_plan(need, want) == want
if rank == 0:
    print("LAYOUT_OK", P, int(sum(need)), int(allc.sum()))
dist.destroy_process_group()
'''


def test_layout_single_rank_is_prefix_sum():
    counts = np.array([[3, 0, 5, 2]], dtype=np.uint64)                      # W = 1, PW = 4: everything is the own chunk
    base, seg = exchange_layout(1, 4, counts, 0)
    assert list(seg[0]) == [0, 3, 3, 8] and list(base) == [0, 0]


def test_layout_two_ranks_by_hand():
    # W = 2, PW = 2 (P = 4, owner(p) = p % 2, q order = [p0, p2 | p1, p3]).  rank0 holds [1,3 | 2,4], rank1 holds [10,30 | 20,40]
    allc = np.array([[1, 3, 2, 4], [10, 30, 20, 40]], dtype=np.uint64)
    b0, s0 = exchange_layout(2, 2, allc, 0)
    b1, s1 = exchange_layout(2, 2, allc, 1)
    # owner 0 receives rank 1's chunk (p0: 10, p2: 30); its own chunk (1, 3) stays local, addressed from 0
    assert list(b0) == [0, 0, 40] and list(s0[0]) == [0, 1] and list(s0[1]) == [0, 10]
    # owner 1 receives rank 0's chunk (p1: 2, p3: 4)
    assert list(b1) == [0, 6, 6] and list(s1[0]) == [0, 2] and list(s1[1]) == [0, 20]


def test_grow_plan_is_monotone():
    assert _grow_plan([10, 20], [0, 0]) == [10 + 2 + 1024, 20 + 5 + 1024]
    assert _grow_plan([10, 20], [100, 5]) == [100, 20 + 5 + 1024]


def test_world_size_2_gloo_agreement(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(script), ROOT]
    p = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    assert "LAYOUT_OK" in p.stdout
