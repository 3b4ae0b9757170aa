#!/usr/bin/env python
"""bench.py — consensus read-pairs/s of the hot path (UMI grouping -> template selection -> score+vote ->
duplex) on synthetic BASELINE.json config 1 ("1M paired 2x150 bp reads, 8 bp prefix UMI, mean cluster
depth 8"), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--impl b200|reference]

A step is one pass of the whole path over one batch (P read pairs per GPU, weak scaling: every rank owns
its own coordinate window).  `value` is measured with the batch resident in HBM (CUDA events, max over
ranks); `e2e` goes through gcb_consensus_batch with pinned HOST buffers, copies inside the timed region;
`roofline` is the score+vote kernel's algorithmic bytes over its own CUDA-event time.
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


def b200_arm(args):
    import torch
    import torch.distributed as dist
    from gencore_b200 import synth
    from gencore_b200.abi import (STAGE_DUPLEX, STAGE_SELECT_TEMPLATE, STAGE_UMI_GROUP, STAGE_VOTE_FAST_ONLY, STAGE_VOTE_PREP_ONLY,
                                  STAGE_VOTE_REST_ONLY, Options)
    from gencore_b200.device import DeviceBatch, DeviceResult, pinned_copy, pinned_result
    from gencore_b200.engine import ConsensusEngine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the consensus engine has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
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
    opt = Options.default()
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

    db = DeviceBatch.from_host(batch, dev)
    dr = DeviceResult.allocate(batch.n_pairs, batch.n_clusters, len(batch.payload) // 4 + 4096, dev)
    # the kernels are launched on this (non-default) stream and the CUDA events are recorded on it
    tstream = torch.cuda.Stream(device=dev)
    stream = tstream.cuda_stream
    assert stream != 0
    torch.cuda.synchronize()
    # the vote is timed in its three parts: per-tile preparation, the ring kernel, rollback + the generic kernel's tiles
    stages = [STAGE_UMI_GROUP, STAGE_SELECT_TEMPLATE, STAGE_VOTE_PREP_ONLY, STAGE_VOTE_FAST_ONLY, STAGE_VOTE_REST_ONLY, STAGE_DUPLEX]
    names = ["umi_group", "select_template+scan", "tile_prep", "score_vote", "rollback+generic", "duplex"]
    i_vote = names.index("score_vote")
    vote_kernel = "vote_ring_kernel"
    vote_parts = [k for k, n in enumerate(names) if n in ("tile_prep", "score_vote", "rollback+generic")]
    def step(events=None):
        for k, st in enumerate(stages):
            if events is not None:
                events[k].record(tstream)
            eng.cluster_by_umi_device(db.struct, dr.struct, st, stream)
        if events is not None:
            events[len(stages)].record(tstream)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    assert eng.batch_status() == 0, "device error flag raised during warm-up"
    res_host = dr.to_host()
    alg = algorithmic_bytes(batch, res_host)

    sampler = ClockSampler(local) if rank == 0 else None
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)] for _ in range(args.steps)]
    launches0 = eng.launches
    barrier()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record(tstream)
    for k in range(args.steps):
        step(evs[k])
    t_end.record(tstream)
    barrier()
    launches = eng.launches - launches0
    total_ms = t_start.elapsed_time(t_end)
    stage_ms = np.zeros(len(stages))
    for k in range(args.steps):
        for s in range(len(stages)):
            stage_ms[s] += evs[k][s].elapsed_time(evs[k][s + 1])
    stage_ms /= args.steps
    if world > 1:
        tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    ms_per_step = total_ms / args.steps
    value = args.pairs * world / (ms_per_step / 1000.0)

    # ---- end to end: host buffers through the C ABI, H2D/D2H inside the timed region
    pb = pinned_copy(batch)
    pr = pinned_result(batch, len(batch.payload) // 4 + 4096)
    for _ in range(2):
        eng.cluster_by_umi(pb, pr)
    barrier()
    if args.host_sweep and rank == 0:  # tuning aid: end-to-end time by kind of host memory for the batch, to stderr
        for label, kw in (("torch pinned", None), ("gcb_host_alloc", dict(lib=eng.lib)), ("gcb_host_alloc write-combined", dict(lib=eng.lib, write_combined=True))):
            pbx = pb if kw is None else pinned_copy(batch, **kw)
            for _ in range(2):
                eng.cluster_by_umi(pbx, pr)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                eng.cluster_by_umi(pbx, pr)
            torch.cuda.synchronize()
            sys.stderr.write("host sweep %s: %.3f ms per step\n" % (label, (time.perf_counter() - t0) / 5 * 1e3))
            del pbx
    if args.chunk_sweep and rank == 0:  # tuning aid: end-to-end time by pipeline chunk size, to stderr
        for mb in [int(x) for x in args.chunk_sweep.split(",")]:
            eng.set_chunk_bytes(mb << 20)
            eng.cluster_by_umi(pb, pr)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                eng.cluster_by_umi(pb, pr)
            torch.cuda.synchronize()
            sys.stderr.write("chunk sweep %d MB: %.3f ms per step\n" % (mb, (time.perf_counter() - t0) / 5 * 1e3))
        eng.set_chunk_bytes(48 << 20)
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        eng.cluster_by_umi(pb, pr)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    clocks = sampler.stop() if sampler else None
    h2d = sum(int(np.ascontiguousarray(getattr(batch, n)).nbytes) for n in ("cluster_pair_off", "cluster_ref", "cluster_flags", "umi", "reads", "cigar", "payload"))
    d2h = int(pr.pair_group.nbytes + pr.cluster_n_groups.nbytes + pr.groups.nbytes + 8 + 4 + int(pr.out_bytes[0]))
    assert np.array_equal(pr.groups, res_host.groups), "host-buffer path and device-buffer path disagree"

    # final gather of per-rank Stats: all counters are additive
    from gencore_b200.hoststats import stats_from_result
    st = stats_from_result(batch, res_host)
    stats_vec = torch.tensor([st.pre_cluster, st.pre_molecule, st.post_sscs, st.post_dcs], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(stats_vec)

    if rank == 0:
        peak, peak_src = peaks()
        vote_ms = float(stage_ms[i_vote])
        whole_ms = float(sum(stage_ms[k] for k in vote_parts))
        # the dominant kernel reads every read's bases and qualities, the reference bases of its slow columns, and writes every
        # consensus record: the whole of SURVEY 8(d)'s algorithmic bytes
        own = dict(alg)
        achieved = own["total"] / (vote_ms * 1e-3) / 1e9
        whole = alg["total"] / (whole_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_gpu": args.pairs, "clusters_per_gpu": batch.n_clusters,
                       "l2": "inputs larger than L2 (payload %d MB per step)" % (len(batch.payload) >> 20),
                       "stage_ms": {n: float(v) for n, v in zip(names, stage_ms)},
                       "stats": {"clusters": int(stats_vec[0]), "molecules": int(stats_vec[1]), "sscs": int(stats_vec[2]), "dcs": int(stats_vec[3])}},
            "roofline": {"bound": "hbm", "kernel": vote_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": args.traffic if args.traffic is not None else (measured_traffic(vote_kernel) if args.pairs == 1_000_000 else None),
                         "algorithmic_bytes": own, "kernel_ms": vote_ms,
                         "timed": "CUDA events around the launch of %s on the launching stream" % vote_kernel,
                         "whole_vote": {"kernels": "every launch of the vote (tile preparation, %s, slow columns, rollback, generic)" % vote_kernel,
                                        "ms": whole_ms, "algorithmic_bytes": alg["total"], "achieved": whole, "frac": whole / peak}},
            "e2e": {"value": args.pairs * world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1000 * e2e_s},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            kind, pairs, per_rep = run_reference_sample(args.cpu_pairs, 1, args.cpu_reps)
            line["cpu_baseline"] = {"value": pairs / float(np.mean(per_rep)), "unit": UNIT, "cores": 1, "kind": kind,
                                    "sample": f"{pairs} pairs of the cfg2 workload x {args.cpu_reps} repetitions, time inside "
                                              f"Cluster::clusterByUMI on one core ({float(np.sum(per_rep)):.1f} s of CPU work)"}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=1_000_000, help="read pairs per GPU per step")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--ref-pairs-per-core", type=int, default=50_000, help="reference arm: pairs per step on each host core")
    ap.add_argument("--cpu-pairs", type=int, default=400_000, help="cpu_baseline sample size (one core)")
    ap.add_argument("--cpu-reps", type=int, default=10, help="cpu_baseline repetitions of the sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
