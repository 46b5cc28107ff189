#!/usr/bin/env python
"""bench.py -- k-mer counting throughput of the B200 hot path (BASELINE.json metric).

  python bench.py --gpus 1 --steps K --warmup W            our arm (CUDA path through the C ABI)
  python bench.py --impl reference --steps K --warmup W    the reference's CPU `dsk` on the host cores

A step = one full pass of the counting path over one synthetic read set:
  value : device-resident leg -- FASTA bytes already in HBM when the clock starts; scan -> super-k-mers ->
          partition (-> exchange over NVLink) -> count -> filter/histogram -> sorted solid set, all on the GPU
          (CUDA events, max over ranks)
  e2e   : same work through the public API with HOST buffers: pinned FASTA bytes in, H2D inside the timed region,
          solid (k-mer, count) set + histogram copied back to the host inside the timed region.

Workloads (config.workload says which):
  N = 1 : BASELINE.json configs[1] -- synthetic 5 Mbp genome, 100x, 150 bp reads, 1 % error, k=31 (400 M k-mers).
  N > 1 : BASELINE.json configs[2] -- ONE synthetic genome of 375 Mbp x N (3 Gbp at N = 8: exactly configs[2]), 30x of
          150 bp reads with 1 % error drawn on the device, every rank parsing its 1/N slice of the reads (9 G k-mers per GPU,
          fixed as N grows: weak scaling; coverage, density and solid share are the same at every N).  72 G k-mers at N = 8.
Inputs (536 MB / 12.2 GB per GPU) are larger than the 126 MB L2, so no explicit flush is needed between iterations.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic HBM bytes per k-mer of the reference dataflow (partition -> LSD radix -> reduce)
A_K = {31: 155.7, 63: 564.4}
C3_GENOME_PER_GPU = 375_000_000          # configs[2] is 3 Gbp over 8 GPUs


def algorithmic_bytes_per_kmer(k, L=150, s=None):
    """SURVEY.md 8(d) formula. W = key bytes, P = digit passes, s = mean k-mers per super-k-mer."""
    W = 8 if k < 32 else 16 if k < 64 else 24 if k < 96 else 32
    P = (2 * k + 7) // 8
    if s is None:
        s = 11.22 if k < 32 else 22.34 if k < 64 else (k - 10 + 2) / 2.0      # ~ half a window of k - m + 1 m-mers (m = 10)
    b_sk = (1 + (k - 1 + s) / 4) / s
    rho = 0.0324 if k < 32 else 0.0337
    s1 = L / (L - k + 1) + b_sk
    s2 = b_sk + W
    s3 = W + 2 * W * P
    s4 = W + (W + 4) * rho
    return {"S1_scan_superk": s1, "S2_expand": s2, "S3_sort": s3, "S4_reduce": s4, "total": s1 + s2 + s3 + s4}


def key_dtype(k):
    return "u64" if k < 32 else "u128" if k < 64 else "u192" if k < 96 else "u256"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_kernels(args):
    """per-kernel figures of the committed `ncu --set full` capture of one bench step at HEAD (profiles/ncu_kernels.json,
    written by tools/ncu_summary.py: ONE capture, one tag, every kernel).  Only valid for the workload it was taken on."""
    p = os.path.join(ROOT, "profiles", "ncu_kernels.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    w = d.get("workload", {})
    if (w.get("kmer_size"), w.get("genome"), w.get("coverage")) != (args.kmer_size, args.genome, args.coverage):
        return None
    return d


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: an NVML polling thread (every ~2 ms, so even a
    60 ms region yields tens of samples); `nvidia-smi -lms 100` only as the fallback when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        import threading
        self.p = None; self.f = None; self.samples = []; self.reasons = set(); self.max_mhz = None; self.power = []
        self._stop = threading.Event(); self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            names = {pynvml.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", pynvml.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}

            def poll():
                while not self._stop.is_set():
                    try:
                        self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        for bit, nm in names.items():
                            if r & bit:
                                self.reasons.add(nm)
                        self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)
                    except Exception:
                        pass
                    time.sleep(0.002)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.thread is not None:
            self._stop.set(); self.thread.join(timeout=2)
            sm = sorted(self.samples)
            if sm:
                out.update(sm_mhz=sm[len(sm) // 2], sm_min_mhz=sm[0], samples=len(sm), source="nvml, 2 ms polling over the timed region")
            if self.power:
                out["power_w_max"] = max(self.power)
            out["sm_max_mhz"] = self.max_mhz
            out["reasons"] = sorted(self.reasons)
            return out
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [ln.strip().split(", ") for ln in open(self.f.name) if ln.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); out["sm_max_mhz"] = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
            out["samples"] = len(sm)
            out["source"] = "nvidia-smi -lms 100"
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------------------ workload
def resolve_workload(args, world):
    """fills the workload defaults for this N and returns the `config` dict BOTH arms print (same keys, same values)"""
    c3 = world > 1 and not args.histo2d
    if args.genome is None:
        args.genome = C3_GENOME_PER_GPU * world if c3 else 5_000_000
    if args.coverage is None:
        args.coverage = 30 if c3 else 100
    if c3:
        args.device_synth = True
    total_reads = int(args.genome * args.coverage // args.read_len)
    reads_per_rank = total_reads // world if c3 else total_reads
    name = "synthetic %.0f Mbp genome, %dx %dbp reads, %.0f%% error, k=%d" % (args.genome / 1e6, args.coverage, args.read_len, args.err * 100, args.kmer_size)
    if (args.genome, args.coverage, args.read_len, args.kmer_size, world) == (5_000_000, 100, 150, 31, 1):
        name += " (BASELINE.json configs[1])"
    if c3:
        name += "; ONE genome, every rank parses its 1/%d slice of the reads" % world
        if args.genome == 8 * C3_GENOME_PER_GPU and world == 8 and args.coverage == 30:
            name += " (BASELINE.json configs[%d])" % (2 if args.kmer_size == 31 else 3)
        else:
            name += " (BASELINE.json configs[2] scaled to %d GPUs: 375 Mbp of genome per GPU, same coverage)" % world
    if args.histo2d:
        name = "-histo2D: %.0f Mbp assembly (bank 0) + %dx %dbp reads, %.0f%% error (bank 1), k=%d (BASELINE.json configs[4] shape)" % (
            args.genome / 1e6, args.coverage, args.read_len, args.err * 100, args.kmer_size)
    cfg = {"workload": name, "kmer_size": args.kmer_size, "genome_bp": int(args.genome), "coverage": args.coverage, "read_len": args.read_len,
           "error_rate": args.err, "seed": args.seed, "reads_total": total_reads if c3 or world == 1 else total_reads * world,
           "abundance_min": 2, "sharding": "reads split over the ranks, partitions owned by p % N" if world > 1 else "one GPU"}
    return cfg, reads_per_rank, c3


def make_workload_host(args, nreads=None, seed_offset=0):
    """FASTA bytes drawn with numpy, in pinned host memory when a device is there"""
    import torch
    from dsk_b200.synth import reads_fasta, genome_codes
    g = genome_codes(args.genome, seed=args.seed)
    n_est = nreads if nreads is not None else int(args.genome * args.coverage // args.read_len)
    est = n_est * (args.read_len + 14) + 1024
    pinned = torch.empty(est, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
    buf = pinned.numpy()
    _, n, nr = reads_fasta(coverage=args.coverage, L=args.read_len, err=args.err, seed=args.seed + seed_offset, out=buf, genome=g, max_reads=nreads)
    return pinned, n, nr


# ------------------------------------------------------------------------------------------------ reference side
def run_reference_dsk(fasta_path, k, cores, tmp, keep=False):
    """one run of the unmodified reference `dsk`; returns (stats text, wall seconds, output prefix)"""
    from oracle.pyoracle import _ref_bin
    out = os.path.join(tmp, "ref_out")
    cmd = [_ref_bin("dsk"), "-file", fasta_path, "-kmer-size", str(k), "-abundance-min", "2", "-histo", "1", "-nb-cores", str(cores),
           "-out", out, "-out-tmp", tmp, "-out-dir", tmp, "-verbose", "1"]
    t0 = time.time()
    p = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True)
    wall = time.time() - t0
    if p.returncode != 0:
        raise RuntimeError("reference dsk failed: " + p.stderr[-500:])
    if not keep:
        for f in os.listdir(tmp):
            if f.startswith("ref_out"):
                try:
                    os.unlink(os.path.join(tmp, f))
                except OSError:
                    pass
    return p.stdout, wall, out


def read_histo(path):
    import numpy as np
    h = np.zeros(10001, np.uint64)
    for line in open(path):
        a, b = line.split()
        h[int(a)] = int(b)
    return h


def dump_arrays(h5_path, k, tmp, tag):
    """dsk2ascii (the reference's own reader) on an .h5 -> sorted (lo, hi, count) arrays"""
    from oracle.pyoracle import _ref_bin, parse_dump
    txt = os.path.join(tmp, tag + ".txt")
    p = subprocess.run([_ref_bin("dsk2ascii"), "-file", h5_path, "-out", txt], cwd=tmp, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("dsk2ascii failed: " + p.stderr[-300:])
    arr = parse_dump(txt, k)
    os.unlink(txt)
    return arr


def digest_arrays(lo, hi, cnt):
    m = hashlib.sha256()
    m.update(lo.tobytes()); m.update(hi.tobytes()); m.update(cnt.tobytes())
    return m.hexdigest()[:16]


def cpu_baseline(args, pinned, n, gpu_result, sample_frac=1.0):
    """times the reference CPU implementation on this box's host cores, on the same FASTA; when the sample is the whole
    workload, also checks the GPU result of the e2e leg against the reference's output VALUE BY VALUE (k-mers, abundances,
    histogram, distinct / solid counts) and times the drop-in CLI (`dsk_gpu`, file -> .h5) next to the reference's.
    Returns (cpu_baseline dict, parity dict or None, cli_e2e dict or None).  Runs after the clock has stopped."""
    import numpy as np
    import oracle
    from oracle.pyoracle import stat_value
    cores = os.cpu_count() or 1
    data = pinned.numpy()[:n]
    if sample_frac < 1.0:
        cut = int(n * sample_frac)
        cut = int(data[:cut].tobytes().rfind(b"\n>")) + 1
        data = data[:cut]
    tmpbase = "/dev/shm" if os.path.isdir("/dev/shm") else None
    if not oracle.ref_available():
        frac = min(sample_frac, 0.05)
        cut = int(n * frac)
        cut = int(data[:cut].tobytes().rfind(b"\n>")) + 1
        t0 = time.time()
        r = oracle.count_files([data[:cut].tobytes()], args.kmer_size, abundance_min=2)
        wall = time.time() - t0
        return ({"value": r.kmers_nb_valid / wall / 1e9, "unit": "Gk-mers/s", "cores": 1, "kind": "port",
                 "sample": "oracle/dsk_oracle.c (scalar port) on %.0f%% of the workload (%d k-mers, %.1f s)" % (100 * frac, r.kmers_nb_valid, wall)}, None, None)
    tmp = tempfile.mkdtemp(prefix="dskbench_", dir=tmpbase)
    parity = None; cli = None
    try:
        fa = os.path.join(tmp, "reads.fa")
        data.tofile(fa)
        full = sample_frac >= 1.0
        stats, wall, out = run_reference_dsk(fa, args.kmer_size, cores, tmp, keep=full)
        nk = int(stat_value(stats, "kmers_nb_valid"))
        base = {"value": nk / wall / 1e9, "unit": "Gk-mers/s", "cores": cores, "kind": "reference",
                "sample": "reference dsk -nb-cores %d -out-tmp tmpfs on %.0f%% of the workload (%d k-mers, %.1f s wall)" % (cores, 100 * sample_frac, nk, wall)}
        if full and gpu_result is not None:
            k = args.kmer_size
            rlo, rhi, rcnt = dump_arrays(out + ".h5", k, tmp, "ref_dump")
            rh = read_histo(out + ".histo")
            keys, cnt, h1, st = gpu_result
            glo = keys[:, 0]; ghi = keys[:, 1] if keys.shape[1] == 2 else np.zeros(len(keys), np.uint64)
            same_n = len(cnt) == len(rcnt)
            parity = {
                "kmers_nb_valid": int(st["kmers_nb_valid"]) == nk,
                "kmers_nb_distinct": int(st["kmers_nb_distinct"]) == int(stat_value(stats, "kmers_nb_distinct")),
                "kmers_nb_solid": int(st["kmers_nb_solid"]) == int(stat_value(stats, "kmers_nb_solid")) == len(rcnt),
                "histogram_identical": bool((h1 == rh).all()),
                "solid_kmers_identical": bool(same_n and (glo == rlo).all() and (ghi == rhi).all()),
                "abundances_identical": bool(same_n and (cnt == rcnt).all()),
                "solid_kmers_compared": int(len(rcnt)),
                "sha256_16_reference": digest_arrays(rlo, rhi, rcnt), "sha256_16_gpu": digest_arrays(np.ascontiguousarray(glo), np.ascontiguousarray(ghi), cnt),
                "how": "unmodified reference dsk on the identical FASTA; its .h5 read back by the reference's dsk2ascii, parsed, sorted, compared value by value with the e2e leg's host result",
            }
            parity["ok"] = all(v for kk, v in parity.items() if isinstance(v, bool))
            # ---- the drop-in CLI: dsk_gpu (host/GpuSortingCount.hpp over the C ABI), file -> .h5, against the reference's CLI
            exe = os.path.join(ROOT, "host", "_build", "dsk_gpu")
            if os.path.exists(exe):
                try:
                    gout = os.path.join(tmp, "gpu_out")
                    cmd = [exe, "-file", fa, "-kmer-size", str(k), "-abundance-min", "2", "-histo", "1", "-out", gout, "-out-dir", tmp, "-verbose", "0"]
                    walls = []
                    for _ in range(2):                                   # second run: CUDA context / module load caches warm
                        t0 = time.time()
                        p = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True)
                        walls.append(time.time() - t0)
                        if p.returncode != 0:
                            raise RuntimeError("dsk_gpu failed: " + (p.stdout + p.stderr)[-400:])
                    clo, chi, ccnt = dump_arrays(gout + ".h5", k, tmp, "gpu_dump")
                    same = len(ccnt) == len(rcnt) and bool((clo == rlo).all() and (chi == rhi).all() and (ccnt == rcnt).all())
                    histo_same = open(gout + ".histo", "rb").read() == open(out + ".histo", "rb").read()
                    cli = {"dsk_gpu_wall_s": min(walls), "dsk_gpu_wall_s_first_run": walls[0], "reference_dsk_wall_s": wall, "speedup": wall / min(walls),
                           "value": nk / min(walls) / 1e9, "unit": "Gk-mers/s", "cores_reference": cores,
                           "h5_dump_identical_to_reference": same, "histo_file_identical_to_reference": histo_same,
                           "what": "wall-clock of the whole CLI, FASTA file on tmpfs -> .h5 + .histo on tmpfs, process start to exit"}
                except Exception as ex:
                    cli = {"error": str(ex)[:300]}
        return base, parity, cli
    finally:
        subprocess.run(["rm", "-rf", tmp])


def reference_arm(args, world, cfg, json_out):
    """--impl reference: the unmodified reference `dsk` (oracle/_ref) on the host cores, same config; a bounded sample of
    the workload per step (the whole workload at N = 1)."""
    import numpy as np
    import oracle
    cores = os.cpu_count() or 1
    total_reads = cfg["reads_total"]
    sample_reads = min(total_reads, int(6e8 // (args.read_len + 12)))           # <= ~600 MB of FASTA per step
    pinned, n, nreads = make_workload_host(args, nreads=sample_reads)
    vals = []
    tmpbase = "/dev/shm" if os.path.isdir("/dev/shm") else None
    if oracle.ref_available():
        from oracle.pyoracle import stat_value
        tmp = tempfile.mkdtemp(prefix="dskbench_", dir=tmpbase)
        fa = os.path.join(tmp, "reads.fa")
        pinned.numpy()[:n].tofile(fa)
        try:
            for i in range(args.warmup + args.steps):
                stats, wall, _ = run_reference_dsk(fa, args.kmer_size, cores, tmp)
                if i >= args.warmup:
                    vals.append((int(stat_value(stats, "kmers_nb_valid")), wall))
        finally:
            subprocess.run(["rm", "-rf", tmp])
        kind = "reference"
        sample = "unmodified reference dsk -nb-cores %d, -out-tmp on tmpfs, %s per step" % (
            cores, "the full workload" if nreads == total_reads else "the first %d of the workload's %d reads (%.2f%%: at this coverage nearly every k-mer is distinct)" % (nreads, total_reads, 100.0 * nreads / total_reads))
    else:
        data = pinned.numpy()[:int(n * 0.05)].tobytes()
        data = data[:data.rfind(b"\n>") + 1]
        for i in range(args.warmup + args.steps):
            t0 = time.time()
            r = oracle.count_files([data], args.kmer_size, abundance_min=2)
            if i >= args.warmup:
                vals.append((r.kmers_nb_valid, time.time() - t0))
        kind, cores, sample = "port", 1, "oracle/dsk_oracle.c scalar port on 5% of the sample per step"
    tot_k = sum(v[0] for v in vals); tot_t = sum(v[1] for v in vals)
    val = tot_k / tot_t / 1e9
    line = {"impl": "reference", "metric": "Gk-mers/s counted", "value": val, "unit": "Gk-mers/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, len(vals)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": key_dtype(args.kmer_size), "data": "synthetic",
            "config": cfg, "kmers_per_step": vals[0][0] if vals else 0,
            "cpu_baseline": {"value": val, "unit": "Gk-mers/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "Gk-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=json_out); json_out.flush()


# ------------------------------------------------------------------------------------------------ multi-GPU parity
def parity_small_job(args, dist, devobj, rank, world, stream):
    """N > 1, before the timed loop: one small seeded job through the SAME exchange + counting path, the union of the ranks'
    solid sets and histograms compared with the CPU oracle on rank 0 (bit-exact).  The oracle is the checker here."""
    import numpy as np
    from dsk_b200 import GpuCounter
    from dsk_b200.distributed import distributed_finish
    from dsk_b200.synth import reads_fasta
    k = args.kmer_size
    buf, n, _ = reads_fasta(G=600_000, coverage=30, L=150, err=0.01, seed=4242 + k)
    data = buf[:n].tobytes()
    cuts = [0]
    for i in range(1, world):
        j = data.find(b"\n>", len(data) * i // world)
        cuts.append(len(data) if j < 0 else j + 1)
    cuts.append(len(data))
    piece = data[cuts[rank]:cuts[rank + 1]]
    eng = GpuCounter(kmer_size=k, abundance_min=2, device=devobj.index, rank=rank, world_size=world, stream=stream.cuda_stream)
    try:
        eng.push_bytes(piece)
        distributed_finish(eng, dist, devobj)
        kk, cc = eng.solid()
        st = eng.stats()
        mine = (kk, cc, eng.histogram()[0], st["kmers_nb_valid"], st["kmers_nb_distinct"])
    finally:
        eng.close()
    allr = [None] * world
    dist.all_gather_object(allr, mine)
    if rank != 0:
        return None
    import oracle
    ref = oracle.count_files([data], k, abundance_min=2)
    keys = np.concatenate([a[0] for a in allr]); cnts = np.concatenate([a[1] for a in allr])
    hist = np.sum([a[2] for a in allr], axis=0, dtype=np.uint64)
    order = np.lexsort((keys[:, 0], keys[:, -1])) if keys.shape[1] == 2 else np.argsort(keys[:, 0], kind="stable")
    keys, cnts = keys[order], cnts[order]
    lo, hi, rc = ref.solid_kmers()
    ok = (sum(a[3] for a in allr) == ref.kmers_nb_valid and sum(a[4] for a in allr) == ref.nb_distinct and len(cnts) == len(rc)
          and bool((keys[:, 0] == lo).all()) and bool((cnts.astype(np.int64) == rc).all()) and bool((hist == ref.hist).all())
          and (keys.shape[1] == 1 or bool((keys[:, 1] == hi).all())))
    return {"ok": bool(ok), "job": "600 kbp x 30x, k=%d, %d ranks: union of the ranks' (k-mer, abundance) sets + summed histogram == oracle" % (k, world),
            "solid_kmers_compared": int(len(rc)), "valid_kmers": int(ref.kmers_nb_valid)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kmer-size", type=int, default=31)
    ap.add_argument("--genome", type=int, default=None, help="default: 5 Mbp at N = 1 (configs[1]), 375 Mbp x N at N > 1 (configs[2])")
    ap.add_argument("--coverage", type=int, default=None, help="default: 100 at N = 1, 30 at N > 1")
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--err", type=float, default=0.01)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--count-mode", default="auto")
    ap.add_argument("--hash-log2-slots", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity-job", action="store_true")
    ap.add_argument("--histo2d", action="store_true", help="BASELINE.json configs[4]: bank 0 = the genome as an assembly, bank 1 = the reads, -histo2D 1")
    ap.add_argument("--device-synth", action="store_true", help="draw the read set on the device (3 Gbp-class workloads)")
    ap.add_argument("--minimizer-size", type=int, default=0, help="0 = what the host adapters do: dskgpu_suggest_minimizer_size(k-mers of the whole job)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    # stdout carries the JSON line and nothing else: whatever libraries print (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    cfg, reads_per_rank, c3 = resolve_workload(args, world if args.impl == "ours" else max(1, args.gpus))

    if args.impl == "reference":
        if rank != 0:
            return 0
        reference_arm(args, max(1, args.gpus), cfg, json_out)
        return 0

    # ------------------------------------------------------------------------------------------ our arm
    import numpy as np
    import torch
    import torch.distributed as dist
    from dsk_b200 import GpuCounter
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    devobj = torch.device("cuda", local)
    # one explicit (non-default) stream carries everything: the library's kernels, torch's copies and the NCCL metadata
    # collectives are ordered on it without host syncs, and the CUDA events below are recorded on the launching stream
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)

    checks = {}
    if world > 1 and not args.no_parity_job:
        checks["parity_small_job"] = parity_small_job(args, dist, devobj, rank, world, stream)

    dev_asm = None
    if args.device_synth:
        from dsk_b200.synth import reads_fasta_device, genome_device, assembly_fasta_device
        gdev = genome_device(args.genome, seed=args.seed, device="cuda")                      # the SAME genome on every rank
        dev, nreads = reads_fasta_device(args.genome, args.coverage, args.read_len, args.err, seed=args.seed + 1000 * rank, device="cuda",
                                         genome=gdev, nreads=reads_per_rank if c3 else None)   # this rank's slice of the reads
        if args.histo2d:
            dev_asm = assembly_fasta_device(gdev)
        del gdev
        n = dev.numel()
        pinned = None
        torch.cuda.empty_cache()
    else:
        pinned, n, nreads = make_workload_host(args, seed_offset=1000 * rank)
        dev = torch.empty(n, dtype=torch.uint8, device="cuda")
        dev.copy_(pinned[:n], non_blocking=False)
        if args.histo2d:
            from dsk_b200.synth import genome_codes, assembly_fasta
            asm = assembly_fasta(genome_codes(args.genome, seed=args.seed))
            dev_asm = torch.frombuffer(bytearray(asm), dtype=torch.uint8).cuda()
    bank_kw = dict(nb_banks=2, per_bank_counts=True, histo2d=True) if args.histo2d else {}
    rb = 1 if args.histo2d else 0                      # bank id of the reads

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # minimizer length from the k-mers of the WHOLE job (all ranks), like host/GpuSortingCount.hpp does from the bank estimate
    from dsk_b200 import _lib as _dsklib
    per_read = max(0, args.read_len - args.kmer_size + 1)
    job_kmers = (cfg["reads_total"] if (c3 or world == 1) else nreads * world) * per_read + (args.genome if args.histo2d else 0)
    msize = args.minimizer_size or _dsklib.lib().dskgpu_suggest_minimizer_size(job_kmers, args.kmer_size)
    bank_kw["minimizer_size"] = msize
    eng = GpuCounter(kmer_size=args.kmer_size, abundance_min=2, device=local, stream=stream.cuda_stream, count_mode=args.count_mode,
                     hash_log2_slots=args.hash_log2_slots, keep_results_on_device=True, rank=rank, world_size=world, **bank_kw)

    from dsk_b200.distributed import distributed_finish

    def step_device():
        eng.reset()
        if dev_asm is not None and rank == 0:             # the assembly enters the job once
            eng.push_device_bytes(dev_asm.data_ptr(), dev_asm.numel(), bank=0, fmt="fasta")
        eng.push_device_bytes(dev.data_ptr(), n, bank=rb, fmt="fasta")
        if world > 1:
            distributed_finish(eng, dist, devobj)
        else:
            eng.finish()

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0; dom_ms = 0.0; dom_n = 0; stage = {}; x_ms = 0.0; x_bytes = 0; push_gap = 0.0; heavy_ms = 0.0; xt_ms = {}
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
        st = eng.stats()
        launches += st["gpu_launches"]; dom_ms += st["ms_dominant_kernel"]; dom_n += st["dominant_kernel_launches"]
        x_ms += st["ms_exchange"]; x_bytes += st["exchange_bytes_out"]
        for kk in ("ms_parse", "ms_superk", "ms_plan", "ms_partition", "ms_count", "ms_sort"):
            stage[kk] = stage.get(kk, 0.0) + st[kk] / args.steps
        push_gap += (st["ms_push_wall"] - st["ms_parse"] - st["ms_superk"]) / args.steps
        heavy_ms += st["ms_count_heavy"] / args.steps
        for kk, v in getattr(eng, "xchg_times", {}).items():
            xt_ms[kk] = xt_ms.get(kk, 0.0) + v / args.steps
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    st = eng.stats()
    kmers = st["kmers_nb_valid"]
    # size-independent invariant of the last timed step (after the clock stopped): every valid k-mer parsed by some rank
    # was counted by the rank owning its partition, i.e. sum_i i * histogram[i] over all ranks == valid k-mers over all
    # ranks (holds while no abundance reaches 10 000, Histogram.hpp:92,221 -- true for the synthetic read sets here)
    h1 = eng.histogram()[0]
    mass = int((h1.astype(np.uint64) * np.arange(h1.size, dtype=np.uint64)).sum())
    t = torch.tensor([ms, float(kmers)], dtype=torch.float64, device="cuda")
    chk = torch.tensor([mass, int(kmers), int(st["kmers_nb_distinct"]), int(st["kmers_nb_solid"])], dtype=torch.int64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dist.all_reduce(chk, op=dist.ReduceOp.SUM)
        ms_all, kmers_all = float(tmax[0]), float(tsum[1])
    else:
        ms_all, kmers_all = ms, float(kmers)
    xt = torch.tensor([x_ms, float(x_bytes)], dtype=torch.float64, device="cuda")
    if world > 1:
        xmax = xt.clone(); dist.all_reduce(xmax, op=dist.ReduceOp.MAX)
        xsum = xt.clone(); dist.all_reduce(xsum, op=dist.ReduceOp.SUM)
        x_ms_max, x_bytes_all = float(xmax[0]), float(xsum[1])
    else:
        x_ms_max, x_bytes_all = 0.0, 0.0
    chk = [int(x) for x in chk.cpu()]
    by_rank = None
    if world > 1:
        by_rank = [None] * world
        dist.all_gather_object(by_rank, {"stage_ms": stage, "exchange_meta_ms": xt_ms, "ms": ms / args.steps, "heavy_ms": heavy_ms})
    checks.update({"histogram_mass": chk[0], "valid_kmers": chk[1], "distinct_kmers": chk[2], "solid_kmers": chk[3],
                   "every_valid_kmer_counted": (chk[0] == chk[1]) if not args.histo2d else None})
    value = kmers_all * args.steps / (ms_all / 1e3) / 1e9
    engine_info = {"count_mode": args.count_mode, "minimizer_size": msize, "sampled_density": st["density_ppm"] / 1e6, "log2_bins": st["log2_bins"],
                   "partitions": int(st["nb_partitions"]), "smem_partitions": int(st["nb_parts_smem"]), "smem_splits": int(st["nb_smem_splits"]),
                   "smem_table_slots": int(st["smem_table_slots"]), "hash_groups": int(st["nb_groups_hash"]), "solid_regrows": int(st["nb_solid_regrows"]),
                   "kmers_per_step_per_gpu": int(kmers), "input_bytes_per_gpu": int(n)}
    eng.close()                                         # its HBM goes back before the e2e leg allocates its own context
    torch.cuda.empty_cache()

    # ---- e2e leg: host buffers in, host results out ------------------------------------------------------
    e2e = None; gpu_result = None
    if not args.no_e2e:
        if pinned is None:
            pinned = torch.empty(n, dtype=torch.uint8, pin_memory=True)
            pinned.copy_(dev)
        h_asm = None
        if dev_asm is not None:
            h_asm = torch.empty(dev_asm.numel(), dtype=torch.uint8, pin_memory=True); h_asm.copy_(dev_asm)
        # what the host fabric gives this rank while every rank copies at once (N = 8: the e2e ceiling is here, not on the GPU)
        barrier()
        hb0, hb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        hb0.record(stream); dev.copy_(pinned[:n], non_blocking=True); hb1.record(stream)
        barrier()
        h2d_gbs = n / (hb0.elapsed_time(hb1) / 1e3) / 1e9
        eng2 = GpuCounter(kmer_size=args.kmer_size, abundance_min=2, device=local, stream=stream.cuda_stream, count_mode=args.count_mode,
                          hash_log2_slots=args.hash_log2_slots, keep_results_on_device=False, rank=rank, world_size=world, **bank_kw)
        hptr = pinned.data_ptr()

        def step_e2e():
            eng2.reset()
            if h_asm is not None and rank == 0:
                eng2.push_bytes((h_asm.data_ptr(), h_asm.numel()), bank=0, fmt="fasta")
            eng2.push_bytes((hptr, n), bank=rb, fmt="fasta")
            if world > 1:
                distributed_finish(eng2, dist, devobj)
            else:
                eng2.finish()
            return eng2.stats()

        for _ in range(max(1, min(args.warmup, 3))):
            st2 = step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            st2 = step_e2e()
        barrier()
        dt = time.perf_counter() - t0
        words = 1 if args.kmer_size < 32 else 2
        d2h = int(st2["kmers_nb_solid"]) * (8 * words + 4) + 10001 * 12 * 8
        tt = torch.tensor([dt, h2d_gbs, -h2d_gbs], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        hsum = torch.tensor([h2d_gbs], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(hsum, op=dist.ReduceOp.SUM)
        e2e = {"value": kmers_all * args.steps / float(tt[0]) / 1e9, "unit": "Gk-mers/s", "h2d_bytes_per_step": int(n) + (h_asm.numel() if h_asm is not None else 0), "d2h_bytes_per_step": d2h,
               "ms_per_step": 1e3 * float(tt[0]) / args.steps,
               "h2d_probe": {"gbs_per_rank_min": -float(tt[2]), "gbs_per_rank_max": float(tt[1]), "gbs_all_ranks": float(hsum[0]),
                             "what": "one pinned -> device copy of the rank's whole input, all ranks at the same time (CUDA events)",
                             "h2d_floor_ms_per_step": 1e3 * n / (-float(tt[2]) * 1e9)}}
        if world == 1 and not args.histo2d:
            kk, cc = eng2.solid()
            gpu_result = (kk, cc, eng2.histogram()[0], st2)
        eng2.close()

    if rank == 0:
        peak, peak_src = peaks()
        ab = algorithmic_bytes_per_kmer(args.kmer_size, args.read_len, s=st["kmers_nb_valid"] / max(1, st["nb_superkmers"]))
        A = A_K.get(args.kmer_size, ab["total"])
        dom_kernel = "k_count_smem" if st["nb_parts_smem"] else ("k_hash_insert" if st["nb_groups_hash"] else "k_rs_onesweep")
        dom_avg_s = dom_ms / 1e3 / max(1, dom_n)
        nk = ncu_kernels(args)
        kd = (nk or {}).get("kernels", {}).get(dom_kernel)
        # roofline of the dominant kernel, honestly: `traffic` = the DRAM bytes it really moves per launch (dram__bytes_read +
        # write, ncu --set full at HEAD, profiles/ncu_kernels.json), `achieved` = that / its live CUDA-event duration, `frac` =
        # achieved / measured HBM peak.  The kernel is NOT HBM-bound (its table never leaves shared memory): `binding` carries the
        # counters that do bind it, from the same capture.  The SURVEY 8(d) figure (k-mers/s x A(k) / peak, a NORMALISED
        # throughput) stays in normalised_throughput_frac and pipeline_roofline.
        traffic = (kd["dram_bytes_per_launch"] if kd else None)
        achieved = (traffic / dom_avg_s / 1e9) if (traffic and dom_avg_s > 0) else None
        alg_per_launch = (ab["S2_expand"] + ab["S3_sort"] + ab["S4_reduce"]) * kmers * args.steps / max(1, dom_n)
        roof = {"bound": "hbm", "kernel": dom_kernel, "launches": int(dom_n), "avg_launch_ms": 1e3 * dom_avg_s,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "peak_source": peak_src,
                "note": "frac is the kernel's real share of HBM bandwidth: small BY DESIGN (the count table lives in shared memory; the LSD-sort dataflow "
                        "A(k) assumes would move algorithmic_bytes_per_launch); what binds the kernel is in `binding`",
                "algorithmic_bytes_per_launch": alg_per_launch,
                "normalised_throughput_frac": (alg_per_launch / dom_avg_s / 1e9 / peak) if dom_avg_s > 0 else None,
                "binding": ({"resource": "integer ALU pipe / SM issue slots", "alu_pipe_pct_of_peak": kd.get("alu_pipe_pct"),
                             "issue_active_pct_of_peak": kd.get("issue_active_pct"), "sm_throughput_pct": kd.get("sm_throughput_pct"),
                             "barrier_stall_per_issue": kd.get("barrier_stall_per_issue"), "frac_of_binding_resource": (kd.get("issue_active_pct") or 0) / 100.0,
                             "source": "profiles/%s (ncu --set full, one gpurun call at HEAD)" % nk.get("tag")} if kd else None)}
        line = {
            "metric": "Gk-mers/s counted", "value": value, "unit": "Gk-mers/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_all / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": key_dtype(args.kmer_size),
            "data": "synthetic (reads drawn on the device)" if args.device_synth else "synthetic",
            "config": cfg,
            "engine": engine_info,
            "l2_policy": "inputs (%.0f MB per GPU) larger than the 126 MB L2; no flush" % (n / 1e6),
            "gpu_launches": int(launches),
            "clocks": clocks,
            "stage_ms": stage,
            "host_ms_per_step": ms_all / args.steps - sum(stage.values()),
            "push_gap_ms_per_step": push_gap, "count_heavy_ms_per_step": heavy_ms,
            "checks": checks,
            "roofline": roof,
            "pipeline_roofline": {"A_k_bytes_per_kmer": A, "achieved": value / world * A, "peak": peak, "unit": "GB/s", "frac": value / world * A / peak,
                                  "peak_nominal": 8000.0, "frac_nominal": value / world * A / 8000.0,
                                  "frac_e2e": (e2e["value"] / world * A / peak) if e2e else None,
                                  "note": "whole step per GPU against SURVEY 8(d) A(k): a NORMALISED throughput (the path moves far fewer HBM bytes than A(k) assumes), not bandwidth utilisation"},
        }
        if nk:
            # every hot kernel of the step: measured DRAM bytes per launch (one ncu capture at HEAD), their bandwidth over the captured
            # launch time against the measured peak, the share of the step (launch list of the same call), the binding counter
            per = {}
            for kn, v in (nk.get("kernels") or {}).items():
                t_us = v.get("captured_launch_us") or 0
                gbs = (v["dram_bytes_per_launch"] / (t_us * 1e-6) / 1e9) if t_us else None
                per[kn] = {"launches_per_step": v.get("launches_per_step"), "share_of_step_pct": v.get("share_of_step_pct"),
                           "dram_bytes_per_launch": v.get("dram_bytes_per_launch"), "dram_gbs": gbs, "hbm_frac": (gbs / peak) if gbs else None,
                           "binding": v.get("binding"), "sm_throughput_pct": v.get("sm_throughput_pct"), "alu_pipe_pct": v.get("alu_pipe_pct"),
                           "issue_active_pct": v.get("issue_active_pct"), "barrier_stall_per_issue": v.get("barrier_stall_per_issue")}
            line["kernels"] = {"source": "profiles/%s" % nk.get("tag"), "per_kernel": per}
        if world > 1:
            # records stored into other ranks' HBM by k_xchg_send (summed over ranks and steps) / the slowest rank's
            # summed copy-kernel time (CUDA events on the context stream), per GPU, against 900 GB/s per direction (nominal) and
            # the 770 GB/s peer copy B200_PROFILING.md measured on this pool
            per_gpu = (x_bytes_all / world) / (x_ms_max / 1e3) / 1e9 if x_ms_max > 0 else 0.0
            line["nvlink"] = {"exchanged_bytes_per_step": x_bytes_all / args.steps, "ms_exchange_per_step": x_ms_max / args.steps,
                              "achieved": per_gpu, "peak": 900.0, "unit": "GB/s per GPU, one direction", "frac": per_gpu / 900.0,
                              "peak_measured_peer_copy": 770.0, "frac_of_measured": per_gpu / 770.0,
                              "share_of_step": (x_ms_max / args.steps) / (ms_all / args.steps),
                              "note": "one contiguous copy per (sender, receiver) pair, 16-byte peer stores; a rank's own partitions never move"}
        if by_rank:
            line["by_rank"] = by_rank
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline and not args.histo2d:
            try:
                base, parity, cli = cpu_baseline(args, pinned, n, gpu_result, sample_frac=min(1.0, 6e8 / max(1, n)))
                line["cpu_baseline"] = base
                if parity is not None:
                    line["checks"]["parity_vs_reference"] = parity
                if cli is not None:
                    line["cli_e2e"] = cli
            except Exception as ex:  # never lose the GPU line to a baseline hiccup
                line["cpu_baseline"] = {"value": None, "unit": "Gk-mers/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % str(ex)[:300]}
        print(json.dumps(line), file=json_out); json_out.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
