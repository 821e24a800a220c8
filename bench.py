#!/usr/bin/env python3
"""bench.py -- reads/s of the HISAT2 alignment hot path on B200 (BASELINE.json).

  python bench.py --gpus N --steps K --warmup W            our CUDA path
  python bench.py --impl reference --gpus N --steps K ...  reference CPU path

Workload (config.workload): BASELINE.json configs[1] -- the linear 22_20-21M
example index, 1M synthetic 101-bp single-end reads per GPU (seeded generator,
tools/simreads.py), --no-spliced-alignment.  One "step" = one pass of the whole
hot path (every read aligned to completion, results appended) over the batch.
  value : whole-job reads/s, inputs already resident in HBM, device-timed
  e2e   : same through ht2gpu_align_batch with pinned HOST buffers (H2D of the
          reads + D2H of the alignment records inside the timed region)
With N>1 (torchrun) every rank aligns its own shard (weak scaling), the only
collective is the start-up NCCL broadcast of the index image.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

DATA = os.path.join(ROOT, "data")
INDEX = os.path.join(DATA, "22_20-21M")
FASTA = os.path.join(DATA, "22_20-21M.fa")
REFBIN = os.path.join(ROOT, "oracle", "_ref", "hisat2-align-s")
RDLEN = 101


def gen_reads(n, seed):
    """(codes uint8 [n,101], names list) -- seeded synthetic reads."""
    import simreads
    _, seq = simreads.load_fasta_codes(FASTA)
    m1, _ = simreads.simulate(seq, n, seed=seed)
    lut = np.zeros(256, dtype=np.uint8)
    for i, ch in enumerate(b"ACGT"):
        lut[ch] = i
    lut[ord("N")] = 4
    return m1, lut[m1]


def seeds_for(codes, names):
    """Vectorised genRandSeed (pat.h:55-91) for FASTA reads (quality 'I')."""
    n, L = codes.shape
    base = np.uint32(((0 + 101) * 59 * 61 * 67 * 71 * 73 * 79 * 83) & 0xffffffff)
    s = np.full(n, base, dtype=np.uint32)
    for i in range(L):
        s ^= (codes[:, i].astype(np.uint32) << np.uint32((i & 15) << 1))
    q = np.uint32(0)
    for i in range(L):
        q ^= np.uint32(ord("I") << ((i & 3) << 3))
    s ^= q
    maxl = max(len(x) for x in names)
    nm = np.zeros((n, maxl), dtype=np.uint32)
    for j, x in enumerate(names):
        nm[j, :len(x)] = np.frombuffer(x, dtype=np.uint8)
    for i in range(maxl):
        s ^= (nm[:, i] << np.uint32((i & 3) << 3))
    return s


class ClockSampler(object):
    def __init__(self, dev):
        self.dev = dev
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def start(self):
        def run():
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
                "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            while not self._stop.is_set():
                try:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=5).stdout.decode().strip()
                    if out:
                        self.rows.append([x.strip() for x in out.split(",")])
                except Exception:
                    pass
                self._stop.wait(0.2)
        self._t = threading.Thread(target=run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = max(int(r[1]) for r in self.rows if r[1].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [names[i] for i in range(4) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons}


def write_fasta(path, ascii_reads, prefix=b"r"):
    n, L = ascii_reads.shape
    # vectorised FASTA writer: ">r<i>\n<seq>\n"
    with open(path, "wb") as f:
        step = 200000
        for s in range(0, n, step):
            e = min(n, s + step)
            out = bytearray()
            blk = ascii_reads[s:e]
            for i in range(e - s):
                out += b">" + prefix + str(s + i).encode() + b"\n"
                out += blk[i].tobytes()
                out += b"\n"
            f.write(out)


def run_reference(fasta, n_reads, threads):
    """Time the unmodified reference binary on a FASTA file; returns seconds (wall)."""
    cmd = [REFBIN, "--no-spliced-alignment", "-f", "-x", INDEX, "-U", fasta, "-S", "/dev/null", "-p", str(threads), "--reorder"]
    t0 = time.time()
    subprocess.run(cmd, check=True, stderr=subprocess.DEVNULL, stdout=subprocess.DEVNULL)
    return time.time() - t0


def pick_threads(sample_fa):
    """The reference's -p scaling collapses on its input/output locks on many-core
    hosts; use the thread count that is fastest on a small sample."""
    ncpu = os.cpu_count() or 1
    cands = sorted(set([c for c in (8, 16, 32, 64, ncpu) if c <= ncpu] + [min(ncpu, 8)]))
    best, bt = cands[0], None
    for c in cands:
        t = run_reference(sample_fa, 0, c)
        if bt is None or t < bt:
            best, bt = c, t
    return best


def reference_arm(args, rank, world):
    if rank != 0:
        return
    line = {"impl": "reference", "metric": "reads_per_sec_aligned", "unit": "reads/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/u32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: 22_20-21M linear index, synthetic 101bp SE reads (bounded sample of --ref-sample reads per step, same generator and seed as the GPU arm), --no-spliced-alignment -k 5"}}
    if not (os.path.exists(REFBIN) and os.path.exists(INDEX + ".1.ht2")):
        line = {"impl": "reference", "unavailable": "oracle/_ref/hisat2-align-s or data/22_20-21M index not present"}
        print(json.dumps(line))
        return
    n = args.ref_sample
    ascii_reads, _ = gen_reads(n, seed=1)
    fa = "/tmp/ht2_bench_ref_%d.fa" % os.getpid()
    write_fasta(fa, ascii_reads)
    small = "/tmp/ht2_bench_ref_small_%d.fa" % os.getpid()
    write_fasta(small, ascii_reads[:50000])
    threads = pick_threads(small)
    for _ in range(args.warmup):
        run_reference(small, 50000, threads)
    times = [run_reference(fa, n, threads) for _ in range(args.steps)]
    os.remove(fa); os.remove(small)
    tot = sum(times)
    v = n * args.steps / tot
    line.update({"value": v, "ms_per_step": 1000.0 * tot / args.steps,
                 "cpu_baseline": {"value": v, "unit": "reads/s", "cores": threads, "kind": "reference",
                                  "sample": "%d reads per step, hisat2-align-s -p %d --reorder, wall clock incl. index load and SAM to /dev/null" % (n, threads)},
                 "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0})
    line["config"]["host_threads"] = threads
    line["config"]["host_cores"] = os.cpu_count()
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--reads", type=int, default=1000000, help="reads per GPU per step")
    ap.add_argument("--ref-sample", type=int, default=500000, help="reads per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--paired", action="store_true", help="BASELINE configs[2] shape: --reads/2 pairs of 2x101 bp per GPU (not the default workload)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return 0

    import torch
    import hisat2_b200 as h2
    from hisat2_b200.parallel import broadcast_image
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the alignment path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # ---- index: parse on rank 0, NCCL-broadcast the packed image, adopt in place
    if world > 1:
        img = h2.Index.build_image(INDEX) if rank == 0 else None
        dev_img = broadcast_image(img, rank, device=torch.device("cuda", local))
        torch.cuda.synchronize()
        prefix = dev_img[:4096].cpu().numpy()
        idx = h2.Index(device_image=(dev_img.data_ptr(), dev_img.numel(), prefix), device=local)
    else:
        idx = h2.Index(INDEX, device=local)
    # ---- synthetic reads for this rank (pinned host memory)
    n = args.reads
    if args.paired:
        import simreads
        _, gseq = simreads.load_fasta_codes(FASTA)
        m1, m2 = simreads.simulate(gseq, n // 2, seed=1 + rank)
        lut = np.zeros(256, dtype=np.uint8)
        for i, ch in enumerate(b"ACGT"):
            lut[ch] = i
        lut[ord("N")] = 4
        codes = np.empty((2 * (n // 2), RDLEN), dtype=np.uint8)
        codes[0::2] = lut[m1]; codes[1::2] = lut[m2]
        n = codes.shape[0]
        base_names = [b"r%d" % (i // 2) for i in range(n)]           # genRandSeed stops at '/'
        names = [b"r%d/%d" % (i // 2, 1 + (i & 1)) for i in range(n)]
        seeds = seeds_for(codes, base_names)
    else:
        _, codes = gen_reads(n, seed=1 + rank)
        names = [b"r%d" % i for i in range(n)]
        seeds = seeds_for(codes, names)
    seq_pin = torch.from_numpy(codes.reshape(-1).copy()).pin_memory()
    offs_pin = torch.arange(0, (n + 1) * RDLEN, RDLEN, dtype=torch.int64).pin_memory()
    seeds_pin = torch.from_numpy(seeds.astype(np.uint32).view(np.int32).copy()).pin_memory()
    batch = h2.ReadBatch(seq_pin.numpy(), offs_pin.numpy().view(np.uint64), seeds_pin.numpy().view(np.uint32), names,
                         paired=args.paired)
    # point the batch at the pinned buffers themselves (ReadBatch may have copied)
    batch.seq = seq_pin.numpy(); batch.offs = offs_pin.numpy().view(np.uint64); batch.seeds = seeds_pin.numpy().view(np.uint32)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    for _ in range(max(args.warmup, 3)):
        r = idx.align(batch)
        r.close()
    # ---- device-resident throughput (value)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    res = idx.align(batch, resident_iters=args.steps)
    kernel_ms = res.ms_kernel
    alg_bytes = int(res.reads["alg_bytes"].astype(np.int64).sum())
    n_lf = int(res.reads["n_lf"].astype(np.int64).sum())
    aligned = int((res.reads["n_aln"] > 0).sum()) if args.paired else int(((res.reads["n_aln"][:, 0] > 0)).sum())
    launches = res.n_launches
    err_reads = int((res.reads["err"] != 0).sum())
    res.close()
    barrier()
    # ---- end to end (e2e): pinned host buffers -> H2D -> kernel -> D2H of results, every step
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    e2e_dev_ms = 0.0
    for _ in range(args.steps):
        r = idx.align(batch)
        h2d, d2h = r.h2d_bytes, r.d2h_bytes
        e2e_dev_ms += r.ms_h2d + r.ms_kernel + r.ms_d2h
        launches += r.n_launches
        r.close()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    barrier()
    # ---- informational: the LF-mapping kernel alone (seed search, ht2gpu_seed_search) on 200k reads of the batch
    lf_map = None
    if rank == 0 and not args.paired:
        k = min(200000, n)
        sub = h2.ReadBatch(batch.seq[:k * RDLEN], batch.offs[:k + 1], batch.seeds[:k], names[:k])
        best = None
        for _ in range(3):
            sr = idx.seed_search(sub, max_range=4)
            if best is None or sr.ms_kernel < best[0]:
                best = (sr.ms_kernel, sr.n_lf, sr.alg_bytes)
            sr.close()
        lf_map = {"kernel": "ht2_seed_kernel (count pass + fill pass)", "reads": k, "ms": best[0], "lf_steps": int(best[1]),
                  "lf_per_s": best[1] / best[0] * 1e3, "achieved": best[2] / best[0] / 1e6, "unit": "GB/s",
                  "note": "algorithmic bytes of ONE pass over the time of BOTH passes; not part of value/e2e"}
    # ---- informational: SAM text for one batch on the host back end (outside the timed regions)
    sam_info = None
    if rank == 0:
        r = idx.align(batch)
        t1 = time.perf_counter()
        txt = idx.format_sam(batch, r)
        sam_info = {"ms_per_step": (time.perf_counter() - t1) * 1000.0, "bytes": len(txt), "host_threads": min(64, os.cpu_count() or 1),
                    "note": "ht2gpu_format_sam (selectByScore, MAPQ, CIGAR/MD:Z, SAM lines) for one batch; not part of value/e2e"}
        del txt
        r.close()
    tk = torch.tensor([kernel_ms, e2e_s * 1000.0], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(alg_bytes), float(n_lf), float(aligned), float(launches), float(err_reads)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tk, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    kernel_ms, e2e_ms = float(tk[0]), float(tk[1])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0
    total_reads = n * world
    value = total_reads * args.steps / (kernel_ms / 1000.0)
    e2e_v = total_reads * args.steps / (e2e_ms / 1000.0)
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    per_launch_bytes = float(tot[0]) / world           # one rank's kernel launch
    per_launch_s = (kernel_ms / 1000.0) / args.steps
    achieved = per_launch_bytes / per_launch_s / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("ht2_align_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    line = {
        "metric": "reads_per_sec_aligned", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": kernel_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/u32", "data": "synthetic",
        "config": {"workload": ("BASELINE configs[2] shape: 22_20-21M linear index, %d synthetic 2x101bp pairs per GPU (--fr, -I 0 -X 1000), --no-spliced-alignment -k 5" % (n // 2)) if args.paired else
                               ("BASELINE configs[1]: 22_20-21M linear index, %d synthetic 101bp SE reads per GPU, --no-spliced-alignment -k 5" % n),
                   "reads_per_gpu": n, "read_len": RDLEN, "parallelism": "read-sharded x%d, index replicated (NCCL broadcast at load)" % world,
                   "l2": "index image 6.4 MB is L2-resident by construction (SURVEY 0.4); read batch (101 MB) + per-thread workspace exceed L2, no flush between steps",
                   "aligned_fraction": float(tot[2]) / total_reads, "lf_steps_per_read": float(tot[1]) / total_reads,
                   "capacity_error_reads": int(tot[4])},
        "e2e": {"value": e2e_v, "unit": "reads/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "device_ms_per_step": e2e_dev_ms / args.steps},
        "sam_backend": sam_info,
        "lf_map": lf_map,
        "gpu_launches": int(tot[3]),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "ht2_align_pool_kernel<8,4>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_kind,
                     "algorithmic_bytes_per_read": float(tot[0]) / total_reads,
                     "note": "algorithmic bytes = sides touched x 32 B (rank sides) + ftab/SA-sample entries + 2-bit reference bytes, counted in-kernel"},
    }
    if dist is not None:
        dist.destroy_process_group()
    # ---- CPU baseline (rank 0, N=1 only): the unmodified reference on a bounded sample
    if world == 1 and not args.no_cpu_baseline and not args.paired and os.path.exists(REFBIN):
        ns = min(args.ref_sample, n)
        ascii_reads, _ = gen_reads(ns, seed=1)
        fa = "/tmp/ht2_bench_cpu_%d.fa" % os.getpid()
        write_fasta(fa, ascii_reads)
        small = fa + ".small"
        write_fasta(small, ascii_reads[:50000])
        threads = pick_threads(small)
        t = run_reference(fa, ns, threads)
        os.remove(fa); os.remove(small)
        line["cpu_baseline"] = {"value": ns / t, "unit": "reads/s", "cores": threads, "kind": "reference",
                                "sample": "%d reads, hisat2-align-s -p %d --reorder (fastest of 8/16/32/64/%d threads on a 50k probe), wall clock incl. index load, SAM to /dev/null"
                                          % (ns, threads, os.cpu_count())}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
