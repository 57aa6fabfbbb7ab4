"""Test-side VCF -> variation graph builder (what `vg construct` + haplotype sampling give the reference's tests,
test/t/50_vg_giraffe.t:10-16): reference FASTA + VCF sites -> node sequences (chopped to 32 bp, as vg's -m 32) and haplotype
walks.  Sites-only VCFs carry no genotypes, so haplotypes are drawn here: the reference walk, the first-alternate walk and
random ones.  Variants that overlap an earlier one are skipped (vg nests them; this builder keeps the graph a DAG of
touching, multi-allelic and multi-node sites, which is what the chain model is tested on)."""
import gzip
from pathlib import Path

import numpy as np

from vg_b200 import synth

GOLD = Path(__file__).parent / "golden" / "vcf"


def read_fasta(path):
    with gzip.open(path, "rt") as f:
        return "".join(l.strip() for l in f if not l.startswith(">")).upper()


def read_vcf(path):
    out = []
    with gzip.open(path, "rt") as f:
        for line in f:
            if line.startswith("#"):
                continue
            c = line.rstrip("\n").split("\t")
            out.append((int(c[1]), c[3].upper(), [a.upper() for a in c[4].split(",")]))
    return out


def build(ref, variants, n_haps=8, seed=1, max_node=32):
    rng = np.random.default_rng(seed)
    node_seqs, items = [], []          # items: ("node", id) | ("site", [[ids...], ...])

    def add(seq):
        ids = []
        for i in range(0, len(seq), max_node):
            node_seqs.append(seq[i:i + max_node]); ids.append(len(node_seqs))
        return ids

    cursor, kept = 0, 0
    for pos, r, alts in sorted(variants):
        start = pos - 1
        if start < cursor or ref[start:start + len(r)] != r or any(set(a) - set("ACGT") for a in alts) or set(r) - set("ACGT"):
            continue
        alleles = [r] + [a for a in alts if a != r]
        # shared prefix / suffix belong to the backbone (vg construct normalises the same way)
        pre = 0
        while all(len(a) > pre for a in alleles) and len({a[pre] for a in alleles}) == 1:
            pre += 1
        suf = 0
        while all(len(a) - pre > suf for a in alleles) and len({a[len(a) - 1 - suf] for a in alleles}) == 1:
            suf += 1
        alleles = [a[pre:len(a) - suf] for a in alleles]
        if len(set(alleles)) < 2:
            continue
        for nid in add(ref[cursor:start + pre]):
            items.append(("node", nid))
        uniq = list(dict.fromkeys(alleles))
        items.append(("site", [add(a) for a in uniq]))
        cursor = start + len(r) - suf
        kept += 1
    for nid in add(ref[cursor:]):
        items.append(("node", nid))
    paths = []
    for h in range(n_haps):
        p = []
        for it in items:
            if it[0] == "node":
                p.append(2 * it[1])
            else:
                n = len(it[1])
                a = 0 if h == 0 else (min(1, n - 1) if h == 1 else (min(2, n - 1) if h == 2 else int(rng.integers(0, n))))
                p += [2 * x for x in it[1][a]]
        paths.append(p)
    used = sorted({v >> 1 for p in paths for v in p})
    remap = {old: new + 1 for new, old in enumerate(used)}
    g = synth.SynthGraph([node_seqs[o - 1] for o in used], [[2 * remap[v >> 1] for v in p] for p in paths], None, slots=None, name="vcf").finish()
    return g, kept


def small_x():
    return build(read_fasta(GOLD / "x.fa.gz"), read_vcf(GOLD / "x.vcf.gz"), n_haps=4, seed=3)


def kgp_100k():
    return build(read_fasta(GOLD / "z_100k.fa.gz"), read_vcf(GOLD / "z_100k.vcf.gz"), n_haps=8, seed=5)
