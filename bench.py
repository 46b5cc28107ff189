#!/usr/bin/env python
"""bench.py -- k-mer counting throughput of the B200 hot path (BASELINE.json metric).

  python bench.py --gpus 1 --steps K --warmup W            our arm (CUDA path through the C ABI)
  python bench.py --impl reference --steps K --warmup W    the reference's CPU `dsk` on the host cores

A step = one full pass of the counting path over one synthetic read set:
  value : device-resident leg -- FASTA bytes already in HBM when the clock starts; scan -> super-k-mers ->
          partition -> count -> filter/histogram -> sorted solid set, all on the GPU (CUDA events, max over ranks)
  e2e   : same work through the public API with HOST buffers: pinned FASTA bytes in, H2D inside the timed region,
          solid (k-mer, count) set + histogram copied back to the host inside the timed region.
Workload at N=1 = BASELINE.json configs[1]: synthetic 5 Mbp genome, 100x, 150 bp reads, 1 % error, k=31.
Inputs (536 MB) are larger than the 126 MB L2, so no explicit flush is needed between iterations.
"""
import argparse
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


def algorithmic_bytes_per_kmer(k, L=150, s=None):
    """SURVEY.md 8(d) formula. W = key bytes, P = digit passes, s = mean k-mers per super-k-mer."""
    W = 8 if k < 32 else 16
    P = (2 * k + 7) // 8
    if s is None:
        s = 11.22 if k < 32 else 22.34
    b_sk = (1 + (k - 1 + s) / 4) / s
    rho = 0.0324 if k < 32 else 0.0337
    s1 = L / (L - k + 1) + b_sk
    s2 = b_sk + W
    s3 = W + 2 * W * P
    s4 = W + (W + 4) * rho
    return {"S1_scan_superk": s1, "S2_expand": s2, "S3_sort": s3, "S4_reduce": s4, "total": s1 + s2 + s3 + s4}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, args):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture (profiles/ncu_traffic.json, written by
    tools/ncu_summary.py); only valid for the workload the capture was taken on (the default one)."""
    default = (args.kmer_size == 31 and args.genome == 5_000_000 and args.coverage == 100 and args.read_len == 150)
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not default or not os.path.exists(p):
        return None
    d = json.load(open(p)).get(kernel)
    if not d or "dram__bytes_read.sum" not in d:
        return None
    return d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]


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


def make_workload(args, rank, world):
    """FASTA bytes of this rank's slice of the read set, in pinned host memory."""
    import numpy as np
    import torch
    from dsk_b200.synth import reads_fasta, genome_codes
    g = genome_codes(args.genome, seed=args.seed)
    # weak scaling: every rank draws its own `coverage`x read set from the same genome (different read seed)
    est = int(args.genome * args.coverage // args.read_len) * (args.read_len + 12) + 1024
    pinned = torch.empty(est, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
    buf = pinned.numpy()
    _, n, nreads = reads_fasta(coverage=args.coverage, L=args.read_len, err=args.err, seed=args.seed + 1000 * rank, out=buf, genome=g)
    return pinned, n, nreads


def cpu_reference_run(fasta_path, k, cores, tmp):
    """one run of the reference `dsk` binary; returns (kmers_nb_valid, seconds of its own `time` stat, wall seconds)"""
    from oracle.pyoracle import _ref_bin, stat_value
    out = os.path.join(tmp, "ref_out")
    cmd = [_ref_bin("dsk"), "-file", fasta_path, "-kmer-size", str(k), "-abundance-min", "2", "-histo", "1", "-nb-cores", str(cores),
           "-out", out, "-out-tmp", tmp, "-out-dir", tmp, "-verbose", "1"]
    t0 = time.time()
    p = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True)
    wall = time.time() - t0
    if p.returncode != 0:
        raise RuntimeError("reference dsk failed: " + p.stderr[-500:])
    nk = int(stat_value(p.stdout, "kmers_nb_valid"))
    for f in os.listdir(tmp):
        if f.startswith("ref_out"):
            try:
                os.unlink(os.path.join(tmp, f))
            except OSError:
                pass
    return nk, wall


def cpu_port_run(data, k):
    import oracle
    t0 = time.time()
    r = oracle.count_files([data], k, abundance_min=2)
    return r.kmers_nb_valid, time.time() - t0


def cpu_baseline(args, pinned, n, sample_frac=1.0):
    """times the reference CPU implementation on this box's host cores; returns dict for the JSON line"""
    import oracle
    cores = os.cpu_count() or 1
    data = pinned.numpy()[:n]
    if sample_frac < 1.0:
        cut = int(n * sample_frac)
        cut = int(data[:cut].tobytes().rfind(b"\n>")) + 1
        data = data[:cut]
    tmpbase = "/dev/shm" if os.path.isdir("/dev/shm") else None
    if oracle.ref_available():
        tmp = tempfile.mkdtemp(prefix="dskbench_", dir=tmpbase)
        fa = os.path.join(tmp, "reads.fa")
        data.tofile(fa)
        try:
            nk, wall = cpu_reference_run(fa, args.kmer_size, cores, tmp)
        finally:
            subprocess.run(["rm", "-rf", tmp])
        return {"value": nk / wall / 1e9, "unit": "Gk-mers/s", "cores": cores, "kind": "reference",
                "sample": "reference dsk -nb-cores %d -out-tmp tmpfs on %.0f%% of the workload (%d k-mers, %.1f s wall)" % (cores, 100 * sample_frac, nk, wall)}
    frac = min(sample_frac, 0.05)
    cut = int(n * frac)
    cut = int(data[:cut].tobytes().rfind(b"\n>")) + 1
    nk, wall = cpu_port_run(data[:cut].tobytes(), args.kmer_size)
    return {"value": nk / wall / 1e9, "unit": "Gk-mers/s", "cores": 1, "kind": "port",
            "sample": "oracle/dsk_oracle.c (scalar port) on %.0f%% of the workload (%d k-mers, %.1f s)" % (100 * frac, nk, wall)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kmer-size", type=int, default=31)
    ap.add_argument("--genome", type=int, default=5_000_000)
    ap.add_argument("--coverage", type=int, default=100)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--err", type=float, default=0.01)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--count-mode", default="auto")
    ap.add_argument("--hash-log2-slots", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--histo2d", action="store_true", help="BASELINE.json configs[4]: bank 0 = the genome as an assembly, bank 1 = the reads, -histo2D 1")
    ap.add_argument("--device-synth", action="store_true", help="draw the read set on the device (3 Gbp-class workloads)")
    ap.add_argument("--minimizer-size", type=int, default=0, help="0 = what the host adapters do: dskgpu_suggest_minimizer_size(k-mers of the whole job)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    # stdout carries the JSON line and nothing else: whatever libraries print (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    workload = "synthetic %.0f Mbp genome, %dx %dbp reads, %.0f%% error, k=%d" % (
        args.genome / 1e6, args.coverage, args.read_len, args.err * 100, args.kmer_size)
    if (args.genome, args.coverage, args.read_len, args.kmer_size) == (5_000_000, 100, 150, 31):
        workload += " (BASELINE.json configs[1])"
    if args.histo2d:
        workload = "-histo2D: %.0f Mbp assembly (bank 0) + %dx %dbp reads, %.0f%% error (bank 1), k=%d (BASELINE.json configs[4] shape)" % (
            args.genome / 1e6, args.coverage, args.read_len, args.err * 100, args.kmer_size)

    import torch
    if args.impl == "reference":
        if rank != 0:
            return 0
        pinned, n, nreads = make_workload(args, 0, 1)
        cores = os.cpu_count() or 1
        import oracle
        vals = []
        tmpbase = "/dev/shm" if os.path.isdir("/dev/shm") else None
        if oracle.ref_available():
            tmp = tempfile.mkdtemp(prefix="dskbench_", dir=tmpbase)
            fa = os.path.join(tmp, "reads.fa")
            pinned.numpy()[:n].tofile(fa)
            try:
                for i in range(args.warmup + args.steps):
                    nk, wall = cpu_reference_run(fa, args.kmer_size, cores, tmp)
                    if i >= args.warmup:
                        vals.append((nk, wall))
            finally:
                subprocess.run(["rm", "-rf", tmp])
            kind, sample = "reference", "unmodified reference dsk -nb-cores %d, -out-tmp on tmpfs, the full workload per step" % cores
        else:
            data = pinned.numpy()[:int(n * 0.05)].tobytes()
            data = data[:data.rfind(b"\n>") + 1]
            for i in range(args.warmup + args.steps):
                nk, wall = cpu_port_run(data, args.kmer_size)
                if i >= args.warmup:
                    vals.append((nk, wall))
            kind, cores, sample = "port", 1, "oracle/dsk_oracle.c scalar port on 5% of the workload per step"
        tot_k = sum(v[0] for v in vals); tot_t = sum(v[1] for v in vals)
        val = tot_k / tot_t / 1e9
        line = {"impl": "reference", "metric": "Gk-mers/s counted", "value": val, "unit": "Gk-mers/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, len(vals)), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": {"workload": workload, "kmers_per_step": vals[0][0] if vals else 0},
                "cpu_baseline": {"value": val, "unit": "Gk-mers/s", "cores": cores, "kind": kind, "sample": sample},
                "e2e": {"value": val, "unit": "Gk-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), file=json_out); json_out.flush()
        return 0

    # ------------------------------------------------------------------------------------------ our arm
    import torch.distributed as dist
    from dsk_b200 import GpuCounter
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev_asm = None
    if args.device_synth:
        from dsk_b200.synth import reads_fasta_device, genome_device, assembly_fasta_device
        gdev = genome_device(args.genome, seed=args.seed, device="cuda")
        dev, nreads = reads_fasta_device(args.genome, args.coverage, args.read_len, args.err, seed=args.seed + 1000 * rank, device="cuda", genome=gdev)
        if args.histo2d:
            dev_asm = assembly_fasta_device(gdev)
        del gdev
        n = dev.numel()
        pinned = None
        if not (args.no_e2e and args.no_cpu_baseline):
            pinned = torch.empty(n, dtype=torch.uint8, pin_memory=True)
            pinned.copy_(dev)
        torch.cuda.empty_cache()
    else:
        pinned, n, nreads = make_workload(args, rank, world)
        dev = torch.empty(n, dtype=torch.uint8, device="cuda")
        dev.copy_(pinned[:n], non_blocking=False)
        if args.histo2d:
            from dsk_b200.synth import genome_codes, assembly_fasta
            asm = assembly_fasta(genome_codes(args.genome, seed=args.seed))
            dev_asm = torch.frombuffer(bytearray(asm), dtype=torch.uint8).cuda()
    h_asm = None
    if dev_asm is not None and not args.no_e2e:
        h_asm = torch.empty(dev_asm.numel(), dtype=torch.uint8, pin_memory=True); h_asm.copy_(dev_asm)
    bank_kw = dict(nb_banks=2, per_bank_counts=True, histo2d=True) if args.histo2d else {}
    rb = 1 if args.histo2d else 0                      # bank id of the reads
    # one explicit (non-default) stream carries everything: the library's kernels, torch's copies and the NCCL metadata
    # collectives are ordered on it without host syncs, and the CUDA events below are recorded on the launching stream
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # minimizer length from the k-mers of the WHOLE job (all ranks), like host/GpuSortingCount.hpp does from the bank estimate
    from dsk_b200 import _lib as _dsklib
    job_kmers = world * (int(args.genome * args.coverage // args.read_len) * max(0, args.read_len - args.kmer_size + 1) + (args.genome if args.histo2d else 0))
    msize = args.minimizer_size or _dsklib.lib().dskgpu_suggest_minimizer_size(job_kmers, args.kmer_size)
    bank_kw["minimizer_size"] = msize
    eng = GpuCounter(kmer_size=args.kmer_size, abundance_min=2, device=local, stream=stream.cuda_stream, count_mode=args.count_mode,
                     hash_log2_slots=args.hash_log2_slots, keep_results_on_device=True, rank=rank, world_size=world, **bank_kw)

    from dsk_b200.distributed import distributed_finish
    devobj = torch.device("cuda", local)

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
    launches = 0; dom_ms = 0.0; dom_n = 0; stage = {}; x_ms = 0.0; x_bytes = 0
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
        st = eng.stats()
        launches += st["gpu_launches"]; dom_ms += st["ms_dominant_kernel"]; dom_n += st["dominant_kernel_launches"]
        x_ms += st["ms_exchange"]; x_bytes += st["exchange_bytes_out"]
        for kk in ("ms_parse", "ms_superk", "ms_partition", "ms_count", "ms_sort"):
            stage[kk] = stage.get(kk, 0.0) + st[kk] / args.steps
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    st = eng.stats()
    kmers = st["kmers_nb_valid"]
    # size-independent invariant of the last timed step (after the clock stopped): every valid k-mer parsed by some rank
    # was counted by the rank owning its partition, i.e. sum_i i * histogram[i] over all ranks == valid k-mers over all
    # ranks (holds while no abundance reaches 10 000, Histogram.hpp:92,221 -- true for the synthetic read sets here)
    import numpy as np
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
    checks = {"histogram_mass": chk[0], "valid_kmers": chk[1], "distinct_kmers": chk[2], "solid_kmers": chk[3],
              "every_valid_kmer_counted": (chk[0] == chk[1]) if not args.histo2d else None}
    value = kmers_all * args.steps / (ms_all / 1e3) / 1e9

    # ---- e2e leg: host buffers in, host results out ------------------------------------------------------
    e2e = None
    if not args.no_e2e:
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

        for _ in range(max(1, args.warmup)):
            st2 = step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            st2 = step_e2e()
        barrier()
        dt = time.perf_counter() - t0
        words = 1 if args.kmer_size < 32 else 2
        d2h = int(st2["kmers_nb_solid"]) * (8 * words + 4) + 10001 * 12 * 8
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": kmers_all * args.steps / float(tt[0]) / 1e9, "unit": "Gk-mers/s", "h2d_bytes_per_step": int(n) + (h_asm.numel() if h_asm is not None else 0), "d2h_bytes_per_step": d2h,
               "ms_per_step": 1e3 * float(tt[0]) / args.steps}
        eng2.close()

    if rank == 0:
        peak, peak_src = peaks()
        ab = algorithmic_bytes_per_kmer(args.kmer_size, args.read_len, s=st["kmers_nb_valid"] / max(1, st["nb_superkmers"]))
        A = A_K.get(args.kmer_size, ab["total"])
        # dominant kernel = the counting kernel (hash insert / radix passes): it does the work of the reference's
        # expand + sort stages (S2 + S3 of SURVEY.md 8(d)); its achieved figure divides that algorithmic volume by
        # its own measured duration.  The hash path legitimately moves far fewer HBM bytes than the LSD dataflow
        # the denominator describes (table traffic stays in L2) -- see DESIGN.md "Roofline accounting".
        dom_bytes_per_launch = (ab["S2_expand"] + ab["S3_sort"]) * kmers * args.steps / max(1, dom_n)
        dom_avg_s = dom_ms / 1e3 / max(1, dom_n)
        achieved = dom_bytes_per_launch / dom_avg_s / 1e9 if dom_avg_s > 0 else 0.0
        dom_kernel = "k_count_smem" if st["nb_parts_smem"] else ("k_hash_insert" if st["nb_groups_hash"] else "k_rs_onesweep")
        line = {
            "metric": "Gk-mers/s counted", "value": value, "unit": "Gk-mers/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_all / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64" if args.kmer_size < 32 else "u128",
            "data": "synthetic",
            "config": {"workload": workload, "kmers_per_step_per_gpu": int(kmers), "input_bytes_per_gpu": int(n), "count_mode": args.count_mode, "minimizer_size": msize, "sampled_density": st["density_ppm"] / 1e6, "log2_bins": st["log2_bins"], "partitions": int(st["nb_partitions"]),
                       "smem_partitions": int(st["nb_parts_smem"]), "smem_splits": int(st["nb_smem_splits"]),
                       "l2_policy": "inputs (%.0f MB) larger than the 126 MB L2; no flush" % (n / 1e6), "parallelism": "1 rank per GPU, partitions sharded by id"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "stage_ms": stage,
            "checks": checks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(dom_kernel, args),
                         "kernel": dom_kernel, "launches": int(dom_n), "avg_launch_ms": 1e3 * dom_avg_s,
                         "peak_source": peak_src, "algorithmic_bytes_per_kmer": ab["S2_expand"] + ab["S3_sort"]},
            "pipeline_roofline": {"A_k_bytes_per_kmer": A, "achieved": value / world * A, "peak": peak, "unit": "GB/s", "frac": value / world * A / peak,
                                  "peak_nominal": 8000.0, "frac_nominal": value / world * A / 8000.0,
                                  "note": "whole step per GPU against SURVEY 8(d) A(k); the hash path moves fewer HBM bytes than A(k) assumes"},
        }
        if world > 1:
            # records stored into other ranks' HBM by k_xchg_copy (summed over ranks and steps) / the slowest rank's
            # summed copy-kernel time (CUDA events on the context stream), per GPU, against 900 GB/s per direction
            per_gpu = (x_bytes_all / world) / (x_ms_max / 1e3) / 1e9 if x_ms_max > 0 else 0.0
            line["nvlink"] = {"exchanged_bytes_per_step": x_bytes_all / args.steps, "ms_exchange_per_step": x_ms_max / args.steps,
                              "achieved": per_gpu, "peak": 900.0, "unit": "GB/s per GPU, one direction", "frac": per_gpu / 900.0,
                              "share_of_step": (x_ms_max / args.steps) / (ms_all / args.steps),
                              "note": "the copy kernel also moves each rank's own partitions (1/N of the records) HBM->HBM inside the same launch"}
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline and not args.histo2d:
            try:
                line["cpu_baseline"] = cpu_baseline(args, pinned, n, sample_frac=min(1.0, 6e8 / max(1, n)))
            except Exception as ex:  # never lose the GPU line to a baseline hiccup
                line["cpu_baseline"] = {"value": None, "unit": "Gk-mers/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % ex}
        print(json.dumps(line), file=json_out); json_out.flush()
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
