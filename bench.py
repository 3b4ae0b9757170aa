#!/usr/bin/env python
"""bench.py — consensus read-pairs/s of the hot path (UMI grouping -> template selection -> score+vote ->
duplex) on synthetic BASELINE.json config 1 ("1M paired 2x150 bp reads, 8 bp prefix UMI, mean cluster
depth 8"), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--impl b200|reference]

A pass is the whole path over one batch (P read pairs per GPU, weak scaling: every rank owns its own
coordinate window); a timed step is `--inner` passes over the resident batch.  `value` is measured with the
batch resident in HBM (CUDA events, max over ranks); `e2e` goes through gcb_consensus_batch with pinned HOST
buffers, copies inside the timed region; `roofline` is the vote kernel's algorithmic bytes over its own
CUDA-event time.  `configs` carries the other BASELINE.json shapes (N = 1), `strong` one input sharded by
contig over the N GPUs; a window of every timed batch is compared with the oracle after the timed region.
`--impl reference` times the reference's own Cluster::clusterByUMI (oracle/_ref, the unmodified reference
sources) on the host cores over a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import dataclasses
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "consensus_read_pairs_per_sec"
UNIT = "read-pairs/s"
WORKLOAD = "cfg2: 2x150 bp pairs, 8 bp prefix UMI, mean cluster depth 8, 1% UMI errors, 10% shared clusters"
SEED = 20261017 + 2


CFG_NAME = "cfg2"


def _cfg(contig_len=None):
    from gencore_b200 import synth
    cfg = synth.CONFIGS[CFG_NAME]
    return dataclasses.replace(cfg, contig_len=contig_len) if contig_len else cfg


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant vote kernel on this workload (profiles/traffic.json,
    written from the ncu --set full capture named there); None when no capture of that kernel is committed."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            d = json.load(f)
        return float(d["traffic"]) if d.get("kernel") == kernel else None
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ reference arm
def _ref_worker(args):
    """One host core: the reference's Cluster::clusterByUMI over its own slice of the sample."""
    idx, n_pairs, reps = args
    from gencore_b200 import synth
    from gencore_b200.abi import Options
    from oracle import pyoracle
    batch, genome, contigs = synth.make_fixed_batch(_cfg(2_000_000), seed=SEED + 7919 * (idx + 1), n_pairs=n_pairs, with_qnames=True)
    if pyoracle.reference_available():
        ref = pyoracle.Reference(Options.default(), batch.umi_prefix, genome, contigs)
        secs = [ref.consensus(batch, want_results=False)[3] for _ in range(reps)]
        ref.close()
        return "reference", secs
    orc = pyoracle.Oracle()
    secs = []
    for _ in range(reps):
        t = time.perf_counter()
        orc.consensus(batch, genome, Options.default())
        secs.append(time.perf_counter() - t)
    return "port", secs


def run_reference_sample(total_pairs: int, cores: int, reps: int):
    import multiprocessing as mp
    per = max(total_pairs // cores, 256)
    if cores == 1:
        kind, secs = _ref_worker((0, per, reps))
        return kind, per, np.asarray(secs)
    with mp.get_context("spawn").Pool(cores) as pool:
        out = pool.map(_ref_worker, [(i, per, reps) for i in range(cores)])
    kind = out[0][0]
    per_rep = np.max(np.asarray([o[1] for o in out]), axis=0)  # a step ends when the slowest core ends
    return kind, per * cores, per_rep


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    kind, pairs, per_rep = run_reference_sample(args.ref_pairs_per_core * cores, cores, args.warmup + args.steps)
    timed = per_rep[args.warmup:]
    value = pairs / float(np.mean(timed))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * float(np.mean(timed)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "impl": "reference",
        "config": {"workload": WORKLOAD, "pairs_per_step": int(pairs), "timed": "inside Cluster::clusterByUMI, slowest core per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{pairs} pairs of the cfg2 workload per step, split over {cores} processes"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for row in self.f.read().strip().splitlines():
            c = [x.strip() for x in row.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------ B200 arm
def algorithmic_bytes(batch, res):
    """payload bytes of every read in, consensus records out, reference slices in (no descriptors)."""
    from gencore_b200.hoststats import group_slots
    l = batch.reads["l_qseq"].astype(np.int64)
    l = l[l >= 0]
    reads_in = int(((l + 1) // 2 + l).sum())
    g = res.groups[group_slots(batch, res)]
    t = g["tmpl_read"].reshape(-1)
    lo = batch.reads["l_qseq"][t[t >= 0]].astype(np.int64)
    out = int(((lo + 1) // 2 + lo).sum())
    ref = int(((lo + 1) // 2).sum())
    return {"reads_in": reads_in, "consensus_out": out, "reference_in": ref, "total": reads_in + out + ref}


STAGE_NAMES = ["umi_group", "select_template+scan", "tile_prep", "score_vote", "rollback+generic", "duplex"]
VOTE_KERNEL = "vote_ring_kernel"


class ShapeRun:
    """One batch of one shape resident on one GPU: device-timed passes (CUDA events on the launching stream, per stage),
    end-to-end passes through gcb_consensus_batch with pinned host buffers, and an oracle check of a window of the batch."""

    def __init__(self, eng, torch, dev, batch, genome):
        from gencore_b200.abi import (STAGE_DUPLEX, STAGE_SELECT_TEMPLATE, STAGE_UMI_GROUP, STAGE_VOTE_FAST_ONLY, STAGE_VOTE_PREP_ONLY,
                                      STAGE_VOTE_REST_ONLY)
        from gencore_b200.device import DeviceBatch, DeviceResult
        self.eng, self.torch, self.dev, self.batch, self.genome = eng, torch, dev, batch, genome
        self.db = DeviceBatch.from_host(batch, dev)
        self.cap = len(batch.payload) // 4 + 4096 if batch.n_pairs >= 100_000 else len(batch.payload)
        self.dr = DeviceResult.allocate(batch.n_pairs, batch.n_clusters, self.cap, dev)
        # the kernels are launched on this (non-default) stream and the CUDA events are recorded on it
        self.tstream = torch.cuda.Stream(device=dev)
        assert self.tstream.cuda_stream != 0
        # the vote is timed in its three parts: per-tile preparation, the ring kernel, rollback + the generic kernel's tiles
        self.stages = [STAGE_UMI_GROUP, STAGE_SELECT_TEMPLATE, STAGE_VOTE_PREP_ONLY, STAGE_VOTE_FAST_ONLY, STAGE_VOTE_REST_ONLY, STAGE_DUPLEX]
        torch.cuda.synchronize()

    def one_pass(self, events=None):
        for k, st in enumerate(self.stages):
            if events is not None:
                events[k].record(self.tstream)
            self.eng.cluster_by_umi_device(self.db.struct, self.dr.struct, st, self.tstream.cuda_stream)
        if events is not None:
            events[len(self.stages)].record(self.tstream)

    def warm(self, n):
        for _ in range(max(n, 3)):
            self.one_pass()
        self.torch.cuda.synchronize()
        assert self.eng.batch_status() == 0, "device error flag raised during warm-up"
        self.res_host = self.dr.to_host()
        self.alg = algorithmic_bytes(self.batch, self.res_host)

    def timed(self, steps, inner, barrier):
        """steps x inner passes between two events, every pass ONE call of the whole path (GCB_STAGE_ALL: what a caller issues);
        then a shorter loop with an event between the stages for the per-stage times.  Returns (total ms, per-pass stage ms)."""
        torch = self.torch
        from gencore_b200.abi import STAGE_ALL
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = self.eng.launches
        t0.record(self.tstream)
        for k in range(steps * inner):
            self.eng.cluster_by_umi_device(self.db.struct, self.dr.struct, STAGE_ALL, self.tstream.cuda_stream)
        t1.record(self.tstream)
        self.timed_launches = self.eng.launches - l0  # kernels launched inside the timed region
        barrier()
        n_stage = max(4, (steps * inner) // 4)
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(self.stages) + 1)] for _ in range(n_stage)]
        for k in range(n_stage):
            self.one_pass(evs[k])
        barrier()
        stage_ms = np.zeros(len(self.stages))
        for e in evs:
            for q in range(len(self.stages)):
                stage_ms[q] += e[q].elapsed_time(e[q + 1])
        return t0.elapsed_time(t1), stage_ms / n_stage

    def e2e(self, steps, barrier):
        """seconds per call of gcb_consensus_batch: pinned host buffers in, results back on the host."""
        from gencore_b200.device import pinned_copy, pinned_result
        self.pb = pinned_copy(self.batch)
        self.pr = pinned_result(self.batch, self.cap)
        for _ in range(2):
            self.eng.cluster_by_umi(self.pb, self.pr)
        barrier()
        t = time.perf_counter()
        for _ in range(steps):
            self.eng.cluster_by_umi(self.pb, self.pr)
        barrier()
        sec = (time.perf_counter() - t) / steps
        assert np.array_equal(self.pr.groups, self.res_host.groups), "host-buffer path and device-buffer path disagree"
        b = self.batch
        self.h2d = sum(int(np.ascontiguousarray(getattr(b, n)).nbytes) for n in ("cluster_pair_off", "cluster_ref", "cluster_flags", "umi", "reads", "cigar", "payload"))
        self.d2h = int(self.pr.pair_group.nbytes + self.pr.cluster_n_groups.nbytes + self.pr.groups.nbytes + 8 + 4 + int(self.pr.out_bytes[0]))
        return sec

    def copy_ceiling(self, steps, barrier):
        """seconds per step of the copies of e2e() ALONE: the same bytes host -> device from the same page-locked buffers on one
        stream, the same bytes device -> host on another, no kernel.  With every rank doing it at once this is the ceiling the
        box's host memory system and PCIe fabric put on `e2e` at this N."""
        torch = self.torch
        ins = [t for t in self.pb._pinned]
        outs = [t for t in self.pr._pinned]
        d_in = [torch.empty(t.numel(), dtype=torch.uint8, device=self.dev) for t in ins]
        used = int(self.pr.out_bytes[0])
        # (the result's record buffer is copied up to the bytes the batch produced, as gcb_consensus_batch does)
        d_out = [torch.empty(t.numel(), dtype=torch.uint8, device=self.dev) for t in outs]
        sizes_out = [min(t.numel(), used) if t.numel() >= self.cap else t.numel() for t in outs]
        s_in, s_out = torch.cuda.Stream(device=self.dev), torch.cuda.Stream(device=self.dev)

        def step():
            with torch.cuda.stream(s_in):
                for h, d in zip(ins, d_in):
                    d.copy_(h, non_blocking=True)
            with torch.cuda.stream(s_out):
                for h, d, n in zip(outs, d_out, sizes_out):
                    h[:n].copy_(d[:n], non_blocking=True)
        for _ in range(2):
            step()
        barrier()
        t = time.perf_counter()
        for _ in range(steps):
            step()
        barrier()
        return (time.perf_counter() - t) / steps

    def check_window(self, opt, n_clusters=3000):
        """After the timed region: a window of the timed batch against the oracle, bit for bit (the checker, never the thing measured)."""
        from gencore_b200.shard import slice_batch
        from gencore_b200.verify import assert_window_equal
        from oracle.pyoracle import Oracle
        b = self.batch
        c0 = max(0, b.n_clusters // 2 - n_clusters // 2)
        c1 = min(b.n_clusters, c0 + n_clusters)
        part = slice_batch(b, c0, c1)
        ref = Oracle().consensus(part, self.genome, opt)
        assert_window_equal(self.res_host, part, ref, c0, int(b.cluster_pair_off[c0]), "bench window")
        return {"clusters": [int(c0), int(c1)], "pairs": int(part.n_pairs), "equal_to_oracle": True}

    def roofline(self, stage_ms, peak):
        vote_ms = float(stage_ms[3])
        whole_ms = float(stage_ms[2] + stage_ms[3] + stage_ms[4])
        # the ring kernel reads every read and writes every record; the reference slices are read by slow_columns_kernel, which
        # belongs to the whole vote
        ring_bytes = self.alg["reads_in"] + self.alg["consensus_out"]
        return vote_ms, whole_ms, ring_bytes / (vote_ms * 1e-3) / 1e9 / peak, self.alg["total"] / (whole_ms * 1e-3) / 1e9 / peak

    def free(self):
        self.db = self.dr = self.pb = self.pr = None


def strong_scaling_leg(args, eng, torch, dist, dev, rank, world, barrier):
    """ONE input sharded by contig.  The cfg3 shape (duplex UMIs, depth 20) on 16 contigs, contig c generated from its
    own seed; rank r owns contigs [16r/N, 16(r+1)/N) and sends each through gcb_consensus_batch (host buffers).  Total work is fixed,
    so value(N) / value(1) is the strong-scaling curve; the digest over every contig's result must be the same for every N."""
    import hashlib
    from gencore_b200 import synth
    from gencore_b200.abi import Options
    from gencore_b200.device import pinned_copy, pinned_result
    NC = 16
    cfg = dataclasses.replace(synth.CONFIGS["cfg3"], n_contigs=1, contig_len=10_000_000)
    per = args.strong_pairs // NC
    mine = [c for c in range(NC) if c * world // NC == rank]
    opt = Options.default()
    digests, items = {}, []
    for c in mine:
        batch, genome, _ = synth.make_batch(cfg, seed=SEED + 31 * (c + 1), n_pairs=per, with_qnames=False)
        items.append((c, pinned_copy(batch), pinned_result(batch, len(batch.payload) // 4 + 4096), genome))
    # every contig's reference slice is set right before its batch (a sharded run keeps only its own contigs' reference on the device)
    def run_all():
        for c, pb, pr, genome in items:
            eng.set_reference(genome)
            eng.cluster_by_umi(pb, pr)
    run_all()
    barrier()
    t = time.perf_counter()
    for _ in range(args.strong_steps):
        run_all()
    barrier()
    sec = (time.perf_counter() - t) / args.strong_steps
    for c, pb, pr, genome in items:
        n = int(pr.out_bytes[0])
        h = hashlib.sha256()
        h.update(pr.cluster_n_groups.tobytes()); h.update(pr.pair_group.tobytes()); h.update(pr.groups.tobytes()); h.update(pr.out_payload[:n].tobytes())
        digests[c] = h.hexdigest()
    if world > 1:
        tt = torch.tensor([sec], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sec = float(tt.item())
        gathered = [None] * world
        dist.all_gather_object(gathered, digests)
        digests = {k: v for d in gathered for k, v in d.items()}
    h = hashlib.sha256()
    for c in range(NC):
        h.update(digests[c].encode())
    return {"workload": "cfg3 shape (2x150, duplex UMI, depth 20), %d pairs on %d contigs, sharded by contig; every contig through "
                        "gcb_consensus_batch with host buffers" % (per * NC, NC),
            "scaling": "strong", "pairs": per * NC, "contigs_per_rank": len(mine), "ms": 1000 * sec, "value": per * NC / sec, "unit": UNIT,
            "result_sha256": h.hexdigest(), "steps": args.strong_steps}


def pin_to_gpu_numa_node(torch, local):
    """Binds this rank's threads (and so its first-touch page-locked staging buffers) to the CPUs of the NUMA node its GPU hangs
    off: with N ranks uploading at once the host's memory system is the shared resource (round 1: 44 GB/s per GPU alone, 16 GB/s
    each at N = 8).  Returns what was done for the JSON line; never fails the run."""
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {"numa_node": None, "note": "the platform reports no NUMA node for the GPU"}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return {"numa_node": node, "note": "none of the node's CPUs is available to this process"}
        os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed)}
    except Exception as e:  # (containers without /sys, older torch: run unpinned)
        return {"numa_node": None, "note": f"not pinned: {type(e).__name__}"}


def bam_to_bam_leg(pairs=300_000, timeout_s=150):
    """Not the metric: the whole tools, sorted BAM in / consensus BAM out, on one synthetic cfg2-shaped BAM (scripts/bam_bench.py in a
    process of its own: the stock reference binary on one core, the reference bound to the C ABI, gencore_b200/bin/gencore_b200, the
    same as two --shard processes plus --merge; outputs compared record by record).  A failure is reported, it does not cost the line."""
    import subprocess
    try:
        p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "bam_bench.py"), str(pairs)], capture_output=True, text=True, timeout=timeout_s)
        if p.returncode != 0:
            return {"error": p.stderr.strip().splitlines()[-1][:300] if p.stderr.strip() else "exit %d" % p.returncode}
        return json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    except Exception as e:  # noqa: BLE001
        return {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}


def b200_arm(args):
    import torch
    import torch.distributed as dist
    from gencore_b200 import synth
    from gencore_b200.abi import Options
    from gencore_b200.engine import ConsensusEngine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the consensus engine has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = pin_to_gpu_numa_node(torch, local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # (NCCL prints its version banner on stdout, which carries the one JSON line)
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cfg = _cfg()
    # every rank builds the same genome (same seed) and its own coordinate window of read pairs
    grng = np.random.Generator(np.random.PCG64(SEED))
    contigs, genome = synth.random_genome(grng, [cfg.contig_len] * cfg.n_contigs)
    batch, _, _ = synth.make_batch(cfg, seed=SEED + 104729 * rank, n_pairs=args.pairs, with_qnames=False, genome_cache=(contigs, genome))
    del contigs
    opt = Options.default(cluster_size_req=cfg.supporting_reads)
    eng = ConsensusEngine(opt, local)
    if args.window_shift:
        eng.set_debug(2, args.window_shift)
    if args.group_lanes:
        eng.set_debug(3, args.group_lanes)
    # the packed reference reaches every GPU by ONE NCCL broadcast from rank 0
    g_dev = torch.from_numpy(genome.packed4).to(dev) if rank == 0 else torch.empty(len(genome.packed4), dtype=torch.uint8, device=dev)
    if world > 1:
        dist.broadcast(g_dev, src=0)
    eng.set_reference_device(g_dev.data_ptr(), g_dev.numel(), genome.contig_off, genome.contig_len, keepalive=g_dev)

    run = ShapeRun(eng, torch, dev, batch, genome)
    run.warm(args.warmup)
    sampler = ClockSampler(local) if rank == 0 else None
    total_ms, stage_ms = run.timed(args.steps, args.inner, barrier)
    launches = run.timed_launches
    clocks = sampler.stop() if sampler else None
    if world > 1:
        tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    ms_per_step = total_ms / args.steps
    pairs_per_step = args.pairs * args.inner
    value = pairs_per_step * world / (ms_per_step / 1000.0)

    # ---- end to end: host buffers through the C ABI, H2D/D2H inside the timed region (one batch per step)
    sampler = ClockSampler(local) if rank == 0 else None
    e2e_s = run.e2e(max(3, args.steps), barrier)
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    clocks_e2e = sampler.stop() if sampler else None
    ceil_s = run.copy_ceiling(max(3, args.steps), barrier)
    if world > 1:
        tt = torch.tensor([ceil_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ceil_s = float(tt.item())
    if args.host_sweep and rank == 0:  # tuning aid: end-to-end time by kind of host memory for the batch, to stderr
        from gencore_b200.device import pinned_copy
        for label, kw in (("torch pinned", None), ("gcb_host_alloc", dict(lib=eng.lib)), ("gcb_host_alloc write-combined", dict(lib=eng.lib, write_combined=True))):
            pbx = run.pb if kw is None else pinned_copy(batch, **kw)
            for _ in range(2):
                eng.cluster_by_umi(pbx, run.pr)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                eng.cluster_by_umi(pbx, run.pr)
            torch.cuda.synchronize()
            sys.stderr.write("host sweep %s: %.3f ms per step\n" % (label, (time.perf_counter() - t0) / 5 * 1e3))
            del pbx
    if args.chunk_sweep and rank == 0:  # tuning aid: end-to-end time by pipeline chunk size, to stderr
        for mb in [int(x) for x in args.chunk_sweep.split(",")]:
            eng.set_chunk_bytes(mb << 20)
            eng.cluster_by_umi(run.pb, run.pr)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                eng.cluster_by_umi(run.pb, run.pr)
            torch.cuda.synchronize()
            sys.stderr.write("chunk sweep %d MB: %.3f ms per step\n" % (mb, (time.perf_counter() - t0) / 5 * 1e3))
        eng.set_chunk_bytes(48 << 20)
    parity = run.check_window(opt) if rank == 0 else None

    # final gather of per-rank Stats: all counters are additive
    from gencore_b200.hoststats import stats_from_result
    st = stats_from_result(batch, run.res_host)
    stats_vec = torch.tensor([st.pre_cluster, st.pre_molecule, st.post_sscs, st.post_dcs], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(stats_vec)
    peak, peak_src = peaks()
    vote_ms, whole_ms, frac, whole_frac = run.roofline(stage_ms, peak)
    alg, h2d, d2h, n_clusters, payload_mb = run.alg, run.h2d, run.d2h, batch.n_clusters, len(batch.payload) >> 20
    run.free()
    del run, batch

    # ---- the other shapes of BASELINE.json, one GPU each: cfg1 at its own size, cfg3-cfg5 as one GPU's share (N = 1 runs only)
    configs = None
    if world == 1 and not args.no_configs:
        configs = {}
        for name in ("cfg1", "cfg3", "cfg4", "cfg5"):
            c = synth.CONFIGS[name]
            n_pairs = min(c.n_pairs, args.pairs)
            c = dataclasses.replace(c, n_contigs=min(c.n_contigs, 2), contig_len=min(c.contig_len, 20_000_000))
            b2, g2, _ = synth.make_batch(c, seed=SEED + 17, n_pairs=n_pairs, with_qnames=False)
            o2 = Options.default(cluster_size_req=c.supporting_reads)
            with ConsensusEngine(o2, local) as e2:
                e2.set_reference(g2)
                r2 = ShapeRun(e2, torch, dev, b2, g2)
                r2.warm(3)
                inner2 = 4 if n_pairs >= 100_000 else 64
                tot, sms = r2.timed(5, inner2, barrier)
                sec = r2.e2e(5, barrier)
                vms, wms, fr, wfr = r2.roofline(sms, peak)
                configs[name] = {"pairs": n_pairs, "clusters": b2.n_clusters, "max_cluster_kb": b2.max_cluster_bytes() >> 10, "options": "-s %d" % c.supporting_reads,
                                 "value": n_pairs * 5 * inner2 / (tot * 1e-3), "e2e": n_pairs / sec, "unit": UNIT,
                                 "stage_ms": {n: float(v) for n, v in zip(STAGE_NAMES, sms)},
                                 "roofline_frac": fr, "whole_vote_frac": wfr, "algorithmic_bytes": r2.alg["total"],
                                 "parity": r2.check_window(o2)}
                r2.free()
            del b2, g2, r2
    strong = strong_scaling_leg(args, eng, torch, dist, dev, rank, world, barrier) if not args.no_strong else None

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_gpu": args.pairs, "passes_per_step": args.inner, "pairs_per_step_per_gpu": pairs_per_step,
                       "clusters_per_gpu": n_clusters,
                       "l2": "inputs larger than L2 (payload %d MB per pass)" % payload_mb,
                       "stage_ms_per_pass": {n: float(v) for n, v in zip(STAGE_NAMES, stage_ms)},
                       "stats": {"clusters": int(stats_vec[0]), "molecules": int(stats_vec[1]), "sscs": int(stats_vec[2]), "dcs": int(stats_vec[3])},
                       "parity": parity},
            "roofline": {"bound": "hbm", "kernel": VOTE_KERNEL, "achieved": frac * peak, "peak": peak, "unit": "GB/s", "frac": frac,
                         "peak_source": peak_src, "traffic": args.traffic if args.traffic is not None else (measured_traffic(VOTE_KERNEL) if args.pairs == 1_000_000 else None),
                         "algorithmic_bytes": {"reads_in": alg["reads_in"], "consensus_out": alg["consensus_out"], "reference_in": 0,
                                               "total": alg["reads_in"] + alg["consensus_out"]}, "kernel_ms": vote_ms,
                         "timed": "CUDA events around every launch of %s on the launching stream, mean over the timed region" % VOTE_KERNEL,
                         "whole_vote": {"kernels": "every launch of the vote (tile_prep2_kernel, %s, slow_columns_kernel, vote_rollback_kernel, score_vote_kernel)" % VOTE_KERNEL,
                                        "ms": whole_ms, "algorithmic_bytes": alg["total"], "achieved": whole_frac * peak, "frac": whole_frac}},
            "e2e": {"value": args.pairs * world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1000 * e2e_s,
                    "pairs_per_step_per_gpu": args.pairs, "steps": max(3, args.steps), "host_affinity": affinity,
                    "copy_ceiling": {"value": args.pairs * world / ceil_s, "unit": UNIT, "ms_per_step": 1000 * ceil_s,
                                     "h2d_gb_per_s_per_gpu": h2d / ceil_s / 1e9, "frac_of_ceiling": ceil_s / e2e_s,
                                     "what": "the same H2D and D2H bytes from the same page-locked buffers on two streams, no kernel, every rank at once"}},
            "gpu_launches": int(launches),
            "clocks": clocks, "clocks_e2e": clocks_e2e,
        }
        if configs is not None:
            line["configs"] = configs
        if strong is not None:
            line["strong"] = strong
        if world == 1 and not args.no_cpu_baseline:
            kind, pairs, per_rep = run_reference_sample(args.cpu_pairs, 1, args.cpu_reps)
            line["cpu_baseline"] = {"value": pairs / float(np.mean(per_rep)), "unit": UNIT, "cores": 1, "kind": kind,
                                    "sample": f"{pairs} pairs of the cfg2 workload x {args.cpu_reps} repetitions, time inside "
                                              f"Cluster::clusterByUMI on one core ({float(np.sum(per_rep)):.1f} s of CPU work)"}
        if world == 1 and not args.no_bam:
            line["bam_to_bam"] = bam_to_bam_leg()
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=1_000_000, help="read pairs per GPU per pass")
    ap.add_argument("--inner", type=int, default=32, help="passes over the resident batch per timed step (a pass takes ~0.4 ms: 32 of them make the "
                                                          "device-timed region of 20 steps a quarter of a second, long enough for the clock sampler)")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` block (cfg1, cfg3, cfg4, cfg5 on one GPU)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg (`strong`)")
    ap.add_argument("--strong-pairs", type=int, default=4_000_000, help="strong-scaling leg: pairs of the one sharded input")
    ap.add_argument("--strong-steps", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--ref-pairs-per-core", type=int, default=50_000, help="reference arm: pairs per step on each host core")
    ap.add_argument("--cpu-pairs", type=int, default=400_000, help="cpu_baseline sample size (one core)")
    ap.add_argument("--cpu-reps", type=int, default=10, help="cpu_baseline repetitions of the sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bam", action="store_true", help="skip the `bam_to_bam` block (the whole tools on one synthetic BAM, N = 1 only)")
    ap.add_argument("--window-shift", type=int, default=0, help="tuning aid: log2 of the vote's tile window (14 or 15; 0 = automatic)")
    ap.add_argument("--group-lanes", type=int, default=0, help="tuning aid: lanes per cluster in umi_group / select_template (8, 16, 32; 0 = automatic)")
    ap.add_argument("--host-sweep", action="store_true", help="tuning aid: end-to-end times by kind of host memory, to stderr")
    ap.add_argument("--chunk-sweep", default="", help="tuning aid: comma-separated pipeline chunk sizes in MB whose end-to-end times go to stderr")
    ap.add_argument("--traffic", type=float, default=None, help="dram bytes per vote launch from an ncu capture (profiles/)")
    ap.add_argument("--config", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"],
                    help="exploration only: another fixed-length shape of BASELINE.json (the contract's line is cfg2, the default)")
    args = ap.parse_args()
    global CFG_NAME, WORKLOAD
    if args.config != "cfg2":
        CFG_NAME = args.config
        WORKLOAD = "%s shape of BASELINE.json (exploration run, not the contract's workload)" % args.config
    if args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()
