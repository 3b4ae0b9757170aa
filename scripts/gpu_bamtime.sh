python - <<'PY'
import dataclasses, os, sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import bamfile
from gencore_b200 import synth
cfg = dataclasses.replace(synth.CONFIGS["cfg2"], contig_len=20_000_000)
batch, genome, contigs = synth.make_fixed_batch(cfg, seed=5, n_pairs=300000, with_qnames=True)
os.makedirs('/tmp/bb', exist_ok=True)
bamfile.genome_to_fasta('/tmp/bb/ref.fa', contigs, genome.names)
print(bamfile.batch_to_bam('/tmp/bb/in.bam', batch, genome))
PY
for i in 1 2; do
( time GCB_TIMING=1 gencore_b200/bin/gencore_b200 -i /tmp/bb/in.bam -o /tmp/bb/my.bam -r /tmp/bb/ref.fa ) 2>&1 | tail -25
done
( time oracle/_ref/gencore -i /tmp/bb/in.bam -o /tmp/bb/ref.bam -r /tmp/bb/ref.fa -j /tmp/bb/r.json -h /tmp/bb/r.html ) 2>&1 | tail -4
