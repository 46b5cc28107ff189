#!/usr/bin/env python
"""Multi-process, multi-GPU parity check of the exchange + counting path (run under torchrun on a box with >= 2 GPUs):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py

Every rank parses its slice of one seeded read set, `distributed_finish` routes the super-k-mer records to the rank
owning their partition over NVLink (CUDA-IPC peer pointers) and every rank counts what it owns.  Rank 0 gathers the
ranks' solid sets + histograms and compares their union with the CPU oracle on the whole read set (bit-exact).
Test infrastructure: the oracle is the checker here, never on the product path.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def split_records(data, parts):
    cuts = [0]
    for i in range(1, parts):
        j = data.find(b"\n>", len(data) * i // parts)
        cuts.append(len(data) if j < 0 else j + 1)
    cuts.append(len(data))
    return [data[cuts[i]:cuts[i + 1]] for i in range(parts)]


def main():
    import torch
    import torch.distributed as dist
    import oracle
    from dsk_b200 import GpuCounter
    from dsk_b200.distributed import distributed_finish
    from dsk_b200.synth import reads_fasta
    rank = int(os.environ.get("RANK", 0)); W = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    import faulthandler
    import time
    t_start = time.time()

    def progress(msg):                               # every rank, so a hang can be localised from the log
        sys.stderr.write("[rank %d +%6.1fs] %s\n" % (rank, time.time() - t_start, msg)); sys.stderr.flush()
    cases = [(31, "auto", 400_000, 30, {}), (63, "auto", 400_000, 30, {}), (31, "hash", 300_000, 20, dict(hash_log2_slots=16)),
             (31, "sort", 300_000, 20, {}), (31, "auto", 300_000, 30, dict(smem_table_slots=256, hash_log2_slots=16)),
             (31, "auto", 2_000_000, 30, {}), (63, "auto", 1_000_000, 30, {})]
    if os.environ.get("MGPU_SHORT"):               # N >= 4 boxes are charged N x: one case per path
        cases = [cases[0], cases[1], cases[4], cases[5]]
    shared = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(shared)
    for ci, (k, mode, G, cov, extra) in enumerate(cases):
        buf, n, _ = reads_fasta(G=G, coverage=cov, L=150, err=0.01, seed=1234 + k)
        data = buf[:n].tobytes()
        piece = split_records(data, W)[rank]
        ref = oracle.count_files([data], k, abundance_min=2) if rank == 0 else None
        faulthandler.dump_traceback_later(120, exit=True)      # a stuck exchange ends the run with stacks instead of burning the box
        # odd cases run on a library-owned stream (explicit host syncs order it against torch's NCCL work), even cases on
        # one shared torch stream (what bench.py does)
        eng = GpuCounter(kmer_size=k, abundance_min=2, device=local, rank=rank, world_size=W, count_mode=mode,
                         stream=(shared.cuda_stream if ci % 2 == 0 else None), **extra)
        for rep in range(2 if os.environ.get("MGPU_SHORT") else 3):                      # second round re-uses receive buffers + peer handles (no IPC re-open)
            progress("case %d k=%d %s rep %d: push" % (ci, k, mode, rep))
            eng.reset()
            eng.push_bytes(piece)
            distributed_finish(eng, dist, dev)
            progress("case %d rep %d: finished, %d partitions" % (ci, rep, eng.stats()["nb_partitions"]))
            kk, cc = eng.solid()
            h1 = eng.histogram()[0]
            st = eng.stats()
            mine = (kk, cc, h1, st["kmers_nb_valid"], st["kmers_nb_distinct"], st["nb_partitions"])
            allr = [None] * W
            dist.all_gather_object(allr, mine)
            if rank == 0:
                keys = np.concatenate([a[0] for a in allr]); cnts = np.concatenate([a[1] for a in allr])
                hist = np.sum([a[2] for a in allr], axis=0, dtype=np.uint64)
                valid = sum(a[3] for a in allr); distinct = sum(a[4] for a in allr)
                order = np.lexsort((keys[:, 0], keys[:, -1])) if keys.shape[1] == 2 else np.argsort(keys[:, 0], kind="stable")
                keys, cnts = keys[order], cnts[order]
                lo, hi, rc = ref.solid_kmers()
                good = (valid == ref.kmers_nb_valid and distinct == ref.nb_distinct and len(cnts) == len(rc)
                        and (keys[:, 0] == lo).all() and (cnts.astype(np.int64) == rc).all() and (hist == ref.hist).all()
                        and (keys.shape[1] == 1 or (keys[:, 1] == hi).all()))
                ok = ok and bool(good)
                print("mgpu_check W=%d k=%d mode=%s G=%d rep=%d: %s  (valid %d, distinct %d, solid %d, partitions %d, per-rank solid %s)" % (
                    W, k, mode, G, rep, "OK" if good else "MISMATCH", valid, distinct, len(cnts), allr[0][5], [len(a[1]) for a in allr]), flush=True)
        faulthandler.cancel_dump_traceback_later()
        eng.close()
    # a rank whose input the device scanner rejects must make EVERY rank raise (nobody is left waiting in a collective),
    # and the job after it must run normally
    buf, n, _ = reads_fasta(G=200_000, coverage=10, L=150, err=0.01, seed=99)
    data = buf[:n].tobytes()
    piece = split_records(data, W)[rank]
    eng = GpuCounter(kmer_size=31, abundance_min=2, device=local, rank=rank, world_size=W, stream=shared.cuda_stream)
    faulthandler.dump_traceback_later(120, exit=True)
    eng.push_bytes(b"@r1\nACGTACGTACGTACGTACGTACGTACGTACGTACGT\nACGTACGT\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\nIIIIIIII\n" if rank == W - 1 else piece)
    raised = False
    try:
        distributed_finish(eng, dist, dev)
    except Exception as ex:                          # noqa: BLE001
        raised = True
        progress("failure case: raised %s" % str(ex)[:80])
    eng.reset()
    eng.push_bytes(piece)
    distributed_finish(eng, dist, dev)
    mass = int((eng.histogram()[0].astype(np.uint64) * np.arange(10001, dtype=np.uint64)).sum())
    t = torch.tensor([1 if raised else 0, mass, int(eng.stats()["kmers_nb_valid"])], dtype=torch.int64, device=dev)
    dist.all_reduce(t)
    faulthandler.cancel_dump_traceback_later()
    eng.close()
    if rank == 0:
        good = int(t[0]) == W and int(t[1]) == int(t[2])
        ok = ok and good
        print("mgpu_check W=%d failure propagation: %s  (%d of %d ranks raised; the next job counted %d of %d k-mers)" % (
            W, "OK" if good else "MISMATCH", int(t[0]), W, int(t[1]), int(t[2])), flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("mgpu_check:", "ALL OK" if ok else "FAILED", flush=True)
    return 0 if int(flag[0]) else 1


if __name__ == "__main__":
    sys.exit(main())
