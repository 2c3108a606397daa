#!/usr/bin/env python
"""bench.py — reads aligned per second of the BASAL hot path on B200 (BASELINE.json metric).

Workload at every N: BASELINE.json configs[1] — `-M A:G` paired-end 2x150 bp GLORI-style reads against a
synthetic 500 Mb reference (24 chromosomes, repeat families, N runs; tools/synth.py, seeds 1002 / 2002).
One "step" = one pass of the whole hot path (read packing + seed scheduling, seed look-up, candidate
verification, pairing, best-hit selection) over one batch of STEP_PAIRS synthetic pairs per GPU.
Reads shard across GPUs (weak scaling: every rank maps its own STEP_PAIRS pairs per step against its own
replica of the index); there is no collective on the data path.

  value      reads/s with the batch already resident in HBM (kernels only, CUDA events on the launching stream)
  e2e        reads/s through the C-ABI call a user makes (bsl_align_pe) with pinned HOST buffers: H2D of the
             bases/offsets and D2H of the result records are inside the timed region
  roofline   verify_candidates (the candidate-verification kernel): algorithmic bytes
             candidates x (4 + 8 (ceil(L/32)+1)) / kernel time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the unmodified reference binary (oracle/_ref/basal -p <all cores>) on a bounded sample

`--impl reference` times the reference binary itself (CPU) on the same config and prints the same line.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import synth  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "basal")
GPU_BIN = os.path.join(ROOT, "basal_b200", "bin", "basal")
CONFIG_ID = int(os.environ.get("BENCH_CONFIG", "2"))
SCALE = float(os.environ.get("BENCH_SCALE", "1.0"))            # <1 only for smoke-testing the script itself
STEP_PAIRS = int(os.environ.get("BENCH_STEP_PAIRS", "1000000"))
METRIC = "reads_aligned_per_sec"
UNIT = "reads/s"


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def workload_name(cfg):
    return (f"configs[{cfg.cid - 1}]: -M {cfg.rule} {'paired-end 2x' if cfg.paired else 'single-end '}{cfg.read_len}bp, "
            f"synthetic {cfg.ref_len / 1e6:.0f} Mb reference, flags {' '.join(cfg.flags) or '(defaults)'} -S 7")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


VERIFY_KERNEL = {
    0: "screen_bits (candidate verification: one-sector XOR/popcount screen on the 1-bit plane + exact masked XOR/popcount count of the survivors)",
    3: "screen_candidates (candidate verification: one-sector masked XOR/popcount screen on the 2-bit planes + exact count of the survivors)",
    4: "verify_candidates (candidate verification with -g: masked XOR/popcount over the whole gathered window)",
}


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
        self.proc = None
        self.gpu = gpu

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                                         stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm / cpu baseline

def scratch_dir():
    """tmpfs when it has room (the reference binary and the GPU binary both read and write their files there), else the default."""
    try:
        st = os.statvfs("/dev/shm")
        if st.f_bavail * st.f_frsize > 24 << 30:
            return "/dev/shm"
    except OSError:
        pass
    return None


def write_sample(cfg, chrs, pairs: int, workdir: str):
    """ref.fa + FASTQ file(s) of `pairs` simulated reads / pairs of the workload (default simulator seed). Returns (n, args)."""
    os.makedirs(workdir, exist_ok=True)
    ref = os.path.join(workdir, "ref.fa")
    if not os.path.exists(ref):
        synth.write_fasta(chrs, ref)
    sim = synth.ReadSimulator(cfg, chrs)
    fa, fb = os.path.join(workdir, "s_1.fq"), os.path.join(workdir, "s_2.fq")
    idx = 0
    first = True
    for m1, m2 in sim.chunks(chunk=500_000, limit=pairs):
        synth.write_fastq(fa, m1, idx, "/1" if cfg.paired else "", append=not first)
        if m2 is not None:
            synth.write_fastq(fb, m2, idx, "/2", append=not first)
        idx += len(m1); first = False
    # file names relative to workdir (the binaries run with cwd=workdir): the reference sprintf()s its whole command line into a
    # 256-byte buffer for the @PG header (main.cpp:410,522) and aborts on long paths
    args = ["-a", "s_1.fq"] + (["-b", "s_2.fq"] if cfg.paired else []) + ["-d", "ref.fa", "-M", cfg.rule] + list(cfg.flags) + ["-S", "7"]
    return idx, args


def run_reference_sample(cfg, chrs, pairs: int, workdir: str, threads: int, index_time: float | None = None, have_files=None):
    """Run the reference binary on `pairs` simulated pairs. Returns (reads, align_seconds, index_seconds); its SAM stays in
    workdir/ref_out.sam."""
    idx, args = have_files if have_files else write_sample(cfg, chrs, pairs, workdir)
    base = [REF_BIN] + args + ["-p", str(threads)]

    def timed(extra):
        t0 = time.perf_counter()
        subprocess.run(base + extra + ["-o", "ref_out.sam"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=workdir)
        return time.perf_counter() - t0
    if index_time is None:
        index_time = timed(["-E", "1"])          # load + pack + seed table only
    total = timed([])
    reads = idx * (2 if cfg.paired else 1)
    return reads, max(total - index_time, 1e-3), index_time


def run_gpu_cli(args, workdir: str, gpus: int = 1):
    """The drop-in `basal` binary of this repo on the same files. Returns (map_seconds, index_seconds, total_seconds):
    the binary's own phase clocks ($BASAL_TIMING), mapping = first batch load to last byte written."""
    env = dict(os.environ); env["BASAL_TIMING"] = "1"; env["BASAL_GPUS"] = str(gpus)
    t0 = time.perf_counter()
    p = subprocess.run([GPU_BIN] + args + ["-p", str(os.cpu_count() or 1), "-o", "gpu_out.sam"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env, cwd=workdir)
    total = time.perf_counter() - t0
    t_map = t_idx = None
    for line in p.stderr.decode(errors="replace").splitlines():
        if line.startswith("[timing]"):
            kv = dict(x.split("=") for x in line.split()[1:])
            t_map = float(kv["map_s"]); t_idx = float(kv["load_ref_s"]) + float(kv["index_s"])
            run_gpu_cli.phases = {k: float(v) for k, v in kv.items()}
    if t_map is None:
        raise RuntimeError("basal did not print its [timing] line")
    return t_map, t_idx, total


def compare_sam(workdir: str):
    """Record-by-record diff of ref_out.sam (reference binary) and gpu_out.sam (this repo's binary): every line except the
    @PG header must be byte-identical; batch completion order differs with -p > 1 (SURVEY trap 1), so records are compared as
    sorted multisets. Returns (records, mismatching lines)."""
    env = dict(os.environ); env["LC_ALL"] = "C"
    outs = []
    for name in ("ref_out.sam", "gpu_out.sam"):
        o = os.path.join(workdir, name + ".sorted")
        with open(o, "wb") as fh:
            g = subprocess.Popen(["grep", "-v", "^@PG", os.path.join(workdir, name)], stdout=subprocess.PIPE)
            subprocess.run(["sort", "-S", "2G", "--parallel=8"], stdin=g.stdout, stdout=fh, env=env, check=True)
            g.wait()
        outs.append(o)
    n = int(subprocess.run(["wc", "-l", outs[0]], capture_output=True, text=True).stdout.split()[0])
    d = subprocess.run(["comm", "-3", outs[0], outs[1]], capture_output=True, env=env)
    bad = [l for l in d.stdout.decode(errors="replace").splitlines() if l.strip()]
    return n, len(bad), bad[:4]


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = synth.baseline_config(CONFIG_ID, SCALE)
    if not os.path.exists(REF_BIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/basal not built (run make -f oracle/Makefile.ref)"}))
        return
    threads = os.cpu_count() or 1
    log(f"reference arm: generating {workload_name(cfg)}")
    chrs = synth.make_reference(cfg)
    work = tempfile.mkdtemp(prefix="bench_ref_", dir=scratch_dir())
    try:
        sample = int(os.environ.get("BENCH_REF_PAIRS", str(min(max(200_000, threads * 50_000), 2_000_000))))
        sample = max(1000, int(sample * min(SCALE * 10, 1.0))) if SCALE < 0.1 else sample
        budget_s = float(os.environ.get("BENCH_REF_BUDGET", "300"))
        t_start = time.perf_counter()
        reads, t_al, t_idx = run_reference_sample(cfg, chrs, sample, work, threads)       # warm-up run, also yields T_index
        per_run = t_al + t_idx
        steps = max(1, min(args.steps, int((budget_s - (time.perf_counter() - t_start)) / max(per_run, 1e-3))))
        times = []
        for _ in range(steps):
            r, t, _ = run_reference_sample(cfg, chrs, sample, work, threads, index_time=t_idx)
            times.append(t)
        tot_t = sum(times)
        value = reads * steps / tot_t
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": 1,
                "ms_per_step": 1000 * tot_t / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
                "data": "synthetic", "config": {"workload": workload_name(cfg), "sample_pairs_per_step": sample,
                                                "timing": "wall-clock of the process minus an index-only (-E 1) run"},
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                                 "sample": f"{sample} pairs/step, {steps} timed process runs, index time {t_idx:.1f}s subtracted"},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
    finally:
        shutil.rmtree(work, ignore_errors=True)


# ------------------------------------------------------------------------------------------------ GPU arm

def bind_rank_to_cpus(local: int, world: int):
    """One process per GPU: keep the rank's caller threads on their own share of the host cores (no migration, and the
    pinned staging buffers they touch first stay on the cores that use them)."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
        if world > 1 and len(cpus) >= world:
            per = len(cpus) // world
            os.sched_setaffinity(0, set(cpus[local * per:(local + 1) * per]))
            return per
        return len(cpus)
    except Exception:
        return os.cpu_count() or 1


def measure(capi, cfg, args, dist, rank, world, local, step_pairs, e2e_reps=3):
    """One workload on this rank's GPU: index build, `value` (batch resident, CUDA events), `e2e` (C-ABI calls from pinned
    host buffers, median of e2e_reps repetitions), roofline inputs, replica check. Returns (line pieces, ctx, chrs)."""
    t0 = time.perf_counter()
    chrs = synth.make_reference(cfg)
    cat, offs, lens = synth.reference_ascii(chrs)
    log(f"rank {rank}: reference generated in {time.perf_counter() - t0:.1f}s")
    params = capi.make_params(rule=cfg.rule, S=7, **_flag_kwargs(cfg))
    ctx = capi.Context(params, device=local)
    t0 = time.perf_counter()
    ctx.index_build(cat, offs, lens)
    t_index = time.perf_counter() - t0
    info = ctx.index_info()
    log(f"rank {rank}: GPU index built in {t_index:.2f}s ({info.n_entries} entries, max_kmer_num {info.max_kmer_num})")
    del cat
    # per-rank read shards: each rank simulates its own pairs (weak scaling)
    sim = synth.ReadSimulator(cfg, chrs)
    sim.rng = np.random.default_rng(2000 + cfg.cid + 7919 * rank)
    n_host_batches = 2
    host = []
    L = cfg.read_len

    def pinned_batch(m, readset, first_index):
        pa = capi.pinned_array(m.size); pa[:] = m.reshape(-1)                       # pinned staging, as a caller that cares about transfer speed would use
        po = capi.pinned_array((len(m) + 1) * 8).view(np.uint64); po[:] = np.arange(len(m) + 1, dtype=np.uint64) * L
        return capi.ReadBatch(pa, po, readset=readset, first_index=first_index)
    for m1, m2 in sim.chunks(chunk=step_pairs, limit=step_pairs * n_host_batches):
        a = pinned_batch(m1, 1 if cfg.paired else 0, len(host) * step_pairs)
        b = pinned_batch(m2, 2, len(host) * step_pairs) if m2 is not None else None
        host.append((a, b))
    n = host[0][0].n
    reads_per_step = n * (2 if cfg.paired else 1)
    # one set of pinned result buffers per in-flight call (a context has six lanes = streams + device buffers; four callers measured
    # best: 196 / 230 / 204 / 224 M reads/s with 3 / 4 / 5 / 6, profiles/README.md)
    N_INFLIGHT = int(os.environ.get("BENCH_INFLIGHT", "4"))
    outs = [(capi.pinned_array(n * capi.HIT_DTYPE.itemsize).view(capi.HIT_DTYPE),
             capi.pinned_array(n * capi.HIT_DTYPE.itemsize).view(capi.HIT_DTYPE),
             capi.pinned_array(n * capi.PAIR_DTYPE.itemsize).view(capi.PAIR_DTYPE)) for _ in range(N_INFLIGHT)]
    oa, ob, op = outs[0]

    def call(i, w=0):
        a, b = host[i % len(host)]
        if b is not None:
            ctx.align_pe(a, b, out=outs[w])
        else:
            ctx.align_se(a, out=outs[w][0])

    def run_e2e(steps):
        """`steps` calls through the C-ABI from N_INFLIGHT caller threads (ctypes drops the GIL inside the call), so that the
        H2D / D2H copies of one batch overlap the kernels of the other, exactly as the `basal` CLI drives a GPU."""
        import threading
        errs = []

        def worker(w):
            try:
                for i in range(w, steps, N_INFLIGHT):
                    call(i, w)
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=worker, args=(w,)) for w in range(N_INFLIGHT)]
        t0 = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        dt = time.perf_counter() - t0
        if errs:
            raise errs[0]
        return dt

    def barrier():
        if dist is not None:
            dist.barrier()

    W = max(args.warmup, 3)
    for i in range(W):
        call(i)
    run_e2e(N_INFLIGHT)          # warms every lane's buffers
    # ---- e2e: K calls through the C-ABI with host buffers (H2D + kernels + D2H inside the timed region); median of the repetitions
    clocks = ClockSampler(local); clocks.start()          # samples every 20 ms across both timed regions (e2e, then kernels only)
    time.sleep(0.5)                                       # nvidia-smi needs a moment before its first sample
    for i in range(2):
        call(i)
    e2e_times = []
    for _ in range(e2e_reps):
        barrier()
        e2e_times.append(run_e2e(args.steps))
    h2d = sum(x.bases.nbytes + x.offsets.nbytes for x in host[0] if x is not None)
    d2h = oa.nbytes + (ob.nbytes + op.nbytes if cfg.paired else 0)
    # ---- value: the same step with the batch resident in HBM (kernels only)
    call(0)
    a0, b0 = host[0]
    for _ in range(2):
        ctx.align_rerun(a0, b0)
    barrier()
    dev_ms = search_ms = pack_ms = pair_ms = lookup_ms = verify_ms = reduce_ms = 0.0
    vbytes = cands = lookups = launches = s_launch = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.align_rerun(a0, b0)
        st = ctx.stats()
        dev_ms += st.ms_device; search_ms += st.ms_search; pack_ms += st.ms_pack; pair_ms += st.ms_pair
        lookup_ms += st.ms_lookup; verify_ms += st.ms_verify; reduce_ms += st.ms_reduce
        vbytes += st.verify_bytes; cands += st.candidates; lookups += st.seed_lookups; launches += st.kernel_launches; s_launch += st.search_launches
    t_wall = time.perf_counter() - t0
    clk = clocks.stop()
    barrier()
    # ---- replicas identical: every rank maps the SAME probe batch against its own replica of the index; the hashes of the
    #      result records must agree (the product's multi-GPU contract: a record depends on the read, its index and the parameters only)
    replica = None
    if dist is not None:
        import hashlib
        import torch
        psim = synth.ReadSimulator(cfg, chrs); psim.rng = np.random.default_rng(424242)
        pm1, pm2 = next(psim.chunks(chunk=20000, limit=20000))
        pa = capi.ReadBatch.from_matrix(pm1, readset=1 if cfg.paired else 0)
        if pm2 is not None:
            ra, rb, rp = ctx.align_pe(pa, capi.ReadBatch.from_matrix(pm2, readset=2))
            blob = ra.tobytes() + rb.tobytes() + rp.tobytes()
        else:
            blob = ctx.align_se(pa).tobytes()
        h = np.frombuffer(hashlib.sha256(blob).digest()[:8], dtype=np.int64).copy()
        mine = torch.tensor(h, device="cuda"); allh = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        replica = {"probe_reads": int(len(pm1) * (2 if pm2 is not None else 1)), "ranks": world, "identical": bool(all(int(x[0]) == int(allh[0][0]) for x in allh))}
    # ---- reduce over ranks: max time, summed work
    t_dev = dev_ms / 1000.0
    e2e_sorted = sorted(e2e_times); t_e2e = e2e_sorted[len(e2e_sorted) // 2]
    if dist is not None:
        import torch
        tt = torch.tensor([t_dev, t_wall] + e2e_times, device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        vals = [float(x) for x in tt.tolist()]
        t_dev, t_wall = vals[0], vals[1]; e2e_times = vals[2:]
        e2e_sorted = sorted(e2e_times); t_e2e = e2e_sorted[len(e2e_sorted) // 2]
    total_reads = reads_per_step * args.steps * world
    value = total_reads / t_dev
    e2e = total_reads / t_e2e
    peak, peak_src = peaks()
    achieved = (vbytes / 1e9) / (verify_ms / 1000.0) if verify_ms > 0 else 0.0
    traffic = None                       # dram__bytes_read+write per launch of the verification kernel, from the committed ncu --set full capture of this workload
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp) and cfg.cid == 2:
        try:
            import hashlib
            tj = json.load(open(tp))
            src = hashlib.sha256(open(os.path.join(ROOT, "basal_b200", "csrc", "align.cu"), "rb").read()).hexdigest()
            # only a capture of the kernels as they are now counts (tools/ncu_traffic.py records the source hash)
            traffic = tj.get("dram_bytes_per_launch") if tj.get("kernel_source_sha256") == src else None
            if traffic is None and rank == 0:
                print("[bench] roofline.traffic: profiles/roofline_traffic.json was captured for another align.cu (tools/ncu_traffic.py re-captures it)", file=sys.stderr)
        except Exception:
            traffic = None
    out = {
        "value": value, "ms_per_step": 1000.0 * t_dev / args.steps,
        "config": {"workload": workload_name(cfg), "pairs_per_step_per_gpu": n, "reads_per_step": reads_per_step * world,
                   "distinct_batches": f"value re-runs ONE resident batch of {n} {'pairs' if cfg.paired else 'reads'} per GPU; e2e alternates {n_host_batches} host batches",
                   "l2": "inputs larger than L2 (index + a 300 MB batch per step); no flush needed",
                   "timing": "CUDA events on the launching stream (first kernel to last kernel), max over ranks",
                   "index_build_s": round(t_index, 2), "parallelism": f"read-sharded x{world}, index replicated, no collective"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h) * world,
                "ms_per_step": 1000.0 * t_e2e / args.steps, "repetitions_reads_per_s": [total_reads / t for t in e2e_times],
                "h2d_gb_per_s_per_gpu": h2d * args.steps / t_e2e / 1e9, "d2h_gb_per_s_per_gpu": d2h * args.steps / t_e2e / 1e9,
                "note": f"bsl_align_pe from pinned host buffers, {N_INFLIGHT} caller threads per GPU (one lane = stream + buffers each), median of {len(e2e_times)} repetitions of {args.steps} steps"},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"bound": "hbm", "kernel": VERIFY_KERNEL.get(cfg.cid, VERIFY_KERNEL[0]), "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
                     "candidates_per_step": cands // max(args.steps, 1), "seed_lookups_per_step": lookups // max(args.steps, 1),
                     "bytes_per_candidate": (vbytes // cands) if cands else None, "launches_per_step": s_launch // max(args.steps, 1),
                     "algorithmic_bytes_per_launch": (vbytes // s_launch) if s_launch else None,
                     "ms_verify_per_step": verify_ms / args.steps, "ms_lookup_per_step": lookup_ms / args.steps, "ms_reduce_per_step": reduce_ms / args.steps,
                     "ms_pack_per_step": pack_ms / args.steps, "ms_pair_per_step": pair_ms / args.steps,
                     "wall_ms_per_step": 1000.0 * t_wall / args.steps},
    }
    if replica is not None:
        out["replica_check"] = replica
    return out, ctx, chrs, W


def gpu_arm(args):
    from basal_b200 import capi
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    cores = bind_rank_to_cpus(local, world)
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = synth.baseline_config(CONFIG_ID, SCALE)
    step_pairs = max(1000, int(STEP_PAIRS * min(1.0, SCALE * 10))) if SCALE < 0.1 else STEP_PAIRS
    m, ctx, chrs, W = measure(capi, cfg, args, dist, rank, world, local, step_pairs)
    line = {"metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": m["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic"}
    for k in ("config", "e2e", "gpu_launches", "clocks", "roofline", "replica_check"):
        if k in m:
            line[k] = m[k]
    line["config"]["host_cores_per_rank"] = cores
    ctx.close()
    # ---- 8 GPUs: BASELINE configs[4] (-M C:T PE 2x150 bp, 3.1 Gb human-scale reference, sharded over the box) rides along in the same line
    if (world >= 8 or os.environ.get("BENCH_C5", "0") == "1") and CONFIG_ID == 2 and SCALE >= 1.0 and os.environ.get("BENCH_NO_C5", "0") != "1":
        try:
            del chrs
            cfg5 = synth.baseline_config(5, SCALE)
            a5 = argparse.Namespace(**vars(args)); a5.steps = max(3, min(args.steps, 5))
            m5, ctx5, chrs5, _ = measure(capi, cfg5, a5, dist, rank, world, local, step_pairs, e2e_reps=1)
            ctx5.close()
            line["config5"] = {"workload": m5["config"]["workload"], "value": m5["value"], "unit": UNIT, "ms_per_step": m5["ms_per_step"], "steps": a5.steps,
                               "e2e": m5["e2e"], "roofline": m5["roofline"], "replica_check": m5.get("replica_check"), "index_build_s": m5["config"]["index_build_s"],
                               "pairs_per_step_per_gpu": m5["config"]["pairs_per_step_per_gpu"],
                               "parity": "full-size parity of this workload: tests/test_gpu_fullsize.py (reference binary, 3.1 Gb; BASAL_SLOW_TESTS=1), evidence in profiles/"}
            chrs = chrs5
        except Exception as e:  # never lose the headline line to the extra block
            line["config5"] = {"error": str(e)[-300:]}
            chrs = None
    parity_failed = False
    if rank == 0:
        if world == 1 and os.environ.get("BENCH_SKIP_CPU", "0") != "1" and os.path.exists(REF_BIN):
            threads = os.cpu_count() or 1
            work = tempfile.mkdtemp(prefix="bench_cpu_", dir=scratch_dir())
            try:
                sample = int(os.environ.get("BENCH_CPU_PAIRS", str(min(max(200_000, threads * 62_500), 1_000_000))))
                if SCALE < 0.1:
                    sample = max(1000, int(sample * SCALE * 10))
                log(f"cpu_baseline: reference binary, {sample} pairs, -p {threads}")
                files = write_sample(cfg, chrs, sample, work)
                r, t, ti = run_reference_sample(cfg, chrs, sample, work, threads, have_files=files)
                line["cpu_baseline"] = {"value": r / t, "unit": UNIT, "cores": threads, "kind": "reference",
                                        "sample": f"{sample} pairs of the same workload, wall-clock minus index-only run ({ti:.1f}s index, {t:.1f}s align)"}
            except Exception as e:  # the baseline is a report, never a reason to lose the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": threads, "kind": "reference", "sample": f"failed: {e}"}
                files = None
            # ---- parity beside the number: the drop-in binary of this repo maps the very same files on the GPU and its SAM is
            #      compared record by record with what the reference binary just printed. `cli_e2e` = reads/s of that process path
            #      (FASTQ text in, SAM text out, index build excluded) on a four times longer input made of the same files.
            try:
                if files is not None and os.path.exists(GPU_BIN):
                    run_gpu_cli(files[1], work)
                    n_rec, n_bad, examples = compare_sam(work)
                    line["parity_check"] = {"reads": r, "sam_records": n_rec, "mismatches": n_bad, "against": "oracle/_ref/basal (unmodified reference) on the same FASTQ files, full-size reference, sorted record-by-record diff"}
                    if n_bad:
                        log("PARITY MISMATCH:", examples)
                        parity_failed = True
                    rep = int(os.environ.get("BENCH_CLI_REPEAT", "4"))
                    big = []
                    for f in [x for x in files[1] if x.endswith(".fq")]:
                        o = f[:-3] + "_x.fq"
                        with open(os.path.join(work, o), "wb") as out:
                            for _ in range(rep):
                                with open(os.path.join(work, f), "rb") as src:
                                    shutil.copyfileobj(src, out, 1 << 24)
                        big.append(o)
                    args_big = [big.pop(0) if x.endswith(".fq") else x for x in files[1]]
                    t_map, t_idx, t_tot = run_gpu_cli(args_big, work)
                    line["cli_e2e"] = {"value": r * rep / t_map, "unit": UNIT, "reads": r * rep, "map_s": t_map, "index_s": t_idx, "process_s": t_tot, "host_threads": threads, "phases": getattr(run_gpu_cli, "phases", None),
                                       "scratch": scratch_dir() or "default tmp dir", "note": "basal_b200/bin/basal: plain FASTQ in, SAM out; mapping phase = first batch load to last byte written (the binary's own clock); phases = seconds summed over the worker threads; the reference binary on these files is cpu_baseline"}
            except Exception as e:
                line["parity_check"] = {"reads": 0, "mismatches": None, "error": str(e)[-300:]}
                parity_failed = True
            finally:
                shutil.rmtree(work, ignore_errors=True)
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    if parity_failed:
        sys.exit(3)


def _flag_kwargs(cfg):
    kw = {}
    f = list(cfg.flags)
    for i in range(0, len(f), 2):
        k, v = f[i], f[i + 1]
        if k == "-v": kw["v"] = v
        elif k == "-g": kw["g"] = int(v)
        elif k == "-s": kw["s"] = int(v); kw["s_given"] = True
        elif k == "-I": kw["I"] = int(v)
        elif k == "-w": kw["w"] = int(v)
    return kw


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
