"""N>1 host logic on CPU: world_size-2 gloo processes agree on the exchange layout the scatter kernel will use."""
import os
import subprocess
import sys

import numpy as np

from dsk_b200.distributed import exchange_layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from dsk_b200.distributed import exchange_layout, all_gather_counts
dist.init_process_group("gloo")
rank, W = dist.get_rank(), dist.get_world_size()
P = 12
rng = np.random.default_rng(100 + rank)
recs = rng.integers(0, 50, P).astype(np.uint64)
recs[rank] = 0                                   # an empty partition somewhere
counts = np.concatenate([recs, recs * 11])
allc = all_gather_counts(dist, counts)           # [W, 2P] -- identical on every rank
off, recv = exchange_layout(W, allc, rank)
# every rank checks the global tiling property with the gathered matrix: for each owner, the slots
# [off_s[p], off_s[p] + allc[s][p]) over senders s and owned partitions p tile [0, recv[o]) exactly once
for o in range(W):
    used = np.zeros(int(recv[o]), dtype=np.int32)
    for s in range(W):
        off_s, recv_s = exchange_layout(W, allc, s)
        assert (recv_s == recv).all()
        for p in range(o, P, W):
            a = int(off_s[p]); used[a:a + int(allc[s][p])] += 1
    assert (used == 1).all(), (o, used)
    assert recv[o] == sum(int(allc[s][p]) for s in range(W) for p in range(o, P, W))
# partitions of one owner appear in increasing id, senders in increasing rank inside a partition
prev = -1
for p in range(rank % W, P, W):
    pass
out = np.concatenate([off, recv])
gathered = all_gather_counts(dist, out)
if rank == 0:
    print("LAYOUT_OK", int(recv.sum()), int(allc[:, :P].sum()))
    assert recv.sum() == allc[:, :P].sum()
dist.destroy_process_group()
'''


def test_layout_single_rank_is_prefix_sum():
    counts = np.array([[3, 0, 5, 2, 30, 0, 50, 20]], dtype=np.uint64)      # P = 4
    off, recv = exchange_layout(1, counts, 0)
    assert list(off) == [0, 3, 3, 8] and list(recv) == [10]


def test_layout_two_ranks_by_hand():
    # P = 4, owner(p) = p % 2.  rank0 holds [1,2,3,4] records, rank1 holds [10,20,30,40]
    allc = np.array([[1, 2, 3, 4, 0, 0, 0, 0], [10, 20, 30, 40, 0, 0, 0, 0]], dtype=np.uint64)
    off0, recv = exchange_layout(2, allc, 0)
    off1, _ = exchange_layout(2, allc, 1)
    # owner 0 receives p0 (1+10) then p2 (3+30); owner 1 receives p1 (2+20) then p3 (4+40)
    assert list(recv) == [44, 66]
    assert list(off0) == [0, 0, 11, 22] and list(off1) == [1, 2, 14, 26]


def test_world_size_2_gloo_agreement(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(script), ROOT]
    p = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    assert "LAYOUT_OK" in p.stdout
