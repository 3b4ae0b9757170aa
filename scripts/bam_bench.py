"""BAM-to-BAM wall time of the host pipeline (gencore_b200/bin/gencore_b200 + libgencore_b200.so on cuda:0) next to the
unmodified reference binary (oracle/_ref/gencore, one core) and to the reference bound to the C ABI (oracle/_ref/gencore_bridged,
integration/gcbbridge.h: the reference's own host code, the engine for Cluster::clusterByUMI) on the same synthetic cfg2-shaped
BAM; checks the outputs match.  Not the bench.py metric (that is the hot path behind the C ABI): this is the whole tool, BGZF included."""
import dataclasses, json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bamfile
from gencore_b200 import build as gbuild, synth
from oracle import pyoracle

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
only_b200 = "--only-b200" in sys.argv  # the tool's own timeline (GCB_TIMING lines on stderr), nothing compared
cfg = dataclasses.replace(synth.CONFIGS["cfg2"], contig_len=20_000_000)
with tempfile.TemporaryDirectory() as td:
    batch, genome, contigs = synth.make_fixed_batch(cfg, seed=20261019, n_pairs=n_pairs, with_qnames=True)
    fa, bam = os.path.join(td, "ref.fa"), os.path.join(td, "in.bam")
    bamfile.genome_to_fasta(fa, contigs, genome.names)
    n_rec = bamfile.batch_to_bam(bam, batch, genome)
    res = {}
    for tag, cmd in (("reference", [pyoracle.REF_BIN, "-i", bam, "-o", os.path.join(td, "ref.bam"), "-r", fa, "-j", os.path.join(td, "r.json"), "-h", os.path.join(td, "r.html")]),
                     ("bridged", [os.path.join(os.path.dirname(pyoracle.REF_BIN), "gencore_bridged"), "-i", bam, "-o", os.path.join(td, "bridged.bam"), "-r", fa,
                                  "-j", os.path.join(td, "b.json"), "-h", os.path.join(td, "b.html")]),
                     ("b200", [gbuild.build_cli(), "-i", bam, "-o", os.path.join(td, "b200.bam"), "-r", fa])):
        if not os.path.exists(cmd[0]) or (only_b200 and tag != "b200"):
            continue
        best = None
        for rep in range(2):
            t = time.perf_counter()
            p = subprocess.run(cmd, capture_output=True, text=True, cwd=td, env=dict(os.environ, GENCORE_B200_ENGINE=gbuild.LIB, GCB_TIMING="1"))
            if tag == "b200":  # (the tool's own per-phase wall times)
                sys.stderr.write(p.stderr)
            dt = time.perf_counter() - t
            assert p.returncode == 0, p.stderr[-1500:]
            best = dt if best is None else min(best, dt)
        res[tag] = best
    if only_b200:
        cmd = [gbuild.build_cli(), "-i", bam, "-o", os.path.join(td, "b200.bam"), "-r", fa]
        for fast in ("0", "1"):
            for rep in range(2):
                t = time.perf_counter()
                p = subprocess.run(cmd, capture_output=True, text=True, cwd=td, env=dict(os.environ, GENCORE_B200_ENGINE=gbuild.LIB, GCB_TIMING="1", GCB_FAST_EXIT=fast))
                dt = time.perf_counter() - t
                sys.stderr.write("== GCB_FAST_EXIT=%s, wall %.3f s\n%s" % (fast, dt, p.stderr))
                res["b200_fast_exit_" + fast] = min(res.get("b200_fast_exit_" + fast, 1e9), dt)
        print(json.dumps(dict(res, pairs=n_pairs)))
        sys.exit(0)
    n_out = bamfile.assert_same_bam(os.path.join(td, "ref.bam"), os.path.join(td, "b200.bam"))
    # the same input over two processes on this GPU (--shard 0/2, 1/2) and the merge: scripts/sharded_bam.py
    sharded = None
    for rep in range(2):
        p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "sharded_bam.py"), "-n", "2", "--gpus", "1", "-i", bam, "-o", os.path.join(td, "sharded.bam"),
                            "-r", fa, "--engine", gbuild.LIB], capture_output=True, text=True, cwd=td)
        assert p.returncode == 0, p.stderr[-1500:]
        line = json.loads(p.stdout.strip().splitlines()[-1])
        sharded = line if sharded is None or line["total_s"] < sharded["total_s"] else sharded
    assert bamfile.assert_same_bam(os.path.join(td, "ref.bam"), os.path.join(td, "sharded.bam")) == n_out
    if "bridged" in res:
        assert bamfile.assert_same_bam(os.path.join(td, "ref.bam"), os.path.join(td, "bridged.bam")) == n_out
    print(json.dumps({"what": "BAM-to-BAM wall time, cfg2 shape", "pairs": n_pairs, "records_in": n_rec, "records_out": n_out, "identical_output": True,
                      "reference_s": res["reference"], "bridged_reference_s": res.get("bridged"), "b200_s": res["b200"],
                      "reference_pairs_per_s": n_pairs / res["reference"], "b200_pairs_per_s": n_pairs / res["b200"], "bam_bytes": os.path.getsize(bam),
                      "b200_two_shards_one_gpu": sharded}))
