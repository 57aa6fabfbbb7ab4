"""Copy the reference's small VCF test inputs into tests/golden/vcf/ (run in the build container, where /root/reference
exists; the tests read only the committed copies):
  x.fa.gz / x.vcf.gz            test/small/x.fa + x.vcf.gz        the graph of test/t/50_vg_giraffe.t
  small.middle.ref.fq           test/reads/small.middle.ref.fq     its 63-bp read (expected score 73, 63 without the bonus)
  z_100k.fa.gz / z_100k.vcf.gz  the first 100 kbp of test/1mb1kgp/z.fa with the z.vcf.gz records in that range
                                (1000 Genomes sites-only: SNPs, indels, multi-allelic sites, touching variants)"""
import gzip, shutil
from pathlib import Path
REF = Path("/root/reference/test")
OUT = Path(__file__).resolve().parents[1] / "tests" / "golden" / "vcf"
OUT.mkdir(parents=True, exist_ok=True)
with open(REF / "small" / "x.fa", "rb") as f, gzip.open(OUT / "x.fa.gz", "wb") as g:
    g.write(f.read())
with gzip.open(REF / "small" / "x.vcf.gz", "rt") as f, gzip.open(OUT / "x.vcf.gz", "wt") as g:
    for line in f:
        if line.startswith("##"):
            continue
        g.write("\t".join(line.rstrip("\n").split("\t")[:5]) + "\n")
shutil.copy(REF / "reads" / "small.middle.ref.fq", OUT / "small.middle.ref.fq")
LIMIT = 100_000
seq = "".join(l.strip() for l in open(REF / "1mb1kgp" / "z.fa") if not l.startswith(">"))[:LIMIT]
with gzip.open(OUT / "z_100k.fa.gz", "wt") as g:
    g.write(">z\n" + seq + "\n")
with gzip.open(REF / "1mb1kgp" / "z.vcf.gz", "rt") as f, gzip.open(OUT / "z_100k.vcf.gz", "wt") as g:
    for line in f:
        if line.startswith("##"):
            continue
        c = line.rstrip("\n").split("\t")[:5]
        if line.startswith("#") or int(c[1]) + len(c[3]) <= LIMIT:
            g.write("\t".join(c) + "\n")
print({p.name: p.stat().st_size for p in OUT.iterdir()})
