"""Seeded synthetic graphs and reads for the BASELINE.json configs.

Everything here is data generation for tests and the benchmark (the reference's
counterparts are `vg construct` + `vg sim`, sim_main.cpp knobs -n -l -s -e -i); none of
it is on the measured path.

Graph family: a chain of *slots*.  A slot is either one backbone node or a bubble whose
alleles are single nodes (an allele may be empty: insertion/deletion sites).  The
per-node distance payload (gb_dist_payload) is exact for this family; tests check it
against brute-force shortest paths.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import capi

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[:] = ord("N")
for a, b in zip(b"ACGT", b"TGCA"):
    _COMP[a] = b


def revcomp_bytes(a: np.ndarray) -> np.ndarray:
    return _COMP[a[..., ::-1]]


@dataclass
class SynthGraph:
    node_seqs: list                 # forward sequence of node id i+1
    paths: list                     # haplotypes as lists of oriented nodes (all forward here)
    dist: np.ndarray                # capi.dist_dt, indexed by node id
    hap_seq: list = field(default_factory=list)      # np.uint8 arrays, one per haplotype
    hap_node: list = field(default_factory=list)     # per base: node id
    hap_off: list = field(default_factory=list)      # per base: offset in node
    slots: list = field(default_factory=list)        # per slot: list of node ids (alleles); 0 = empty allele
    name: str = ""

    def build_index(self, k=29, w=11) -> capi.HostIndex:
        return capi.HostIndex(self.node_seqs, self.paths, self.dist, k=k, w=w)

    def finish(self):
        """Derive haplotype sequences and base->(node, offset) maps from the paths."""
        lens = np.array([0] + [len(s) for s in self.node_seqs], dtype=np.int64)
        seq_arrays = [None] + [np.frombuffer(s.encode(), dtype=np.uint8) for s in self.node_seqs]
        self.hap_seq, self.hap_node, self.hap_off = [], [], []
        for p in self.paths:
            ids = np.asarray(p, dtype=np.int64) >> 1
            self.hap_seq.append(np.concatenate([seq_arrays[i] for i in ids]))
            l = lens[ids]
            self.hap_node.append(np.repeat(ids, l).astype(np.uint32))
            starts = np.cumsum(l) - l
            self.hap_off.append((np.arange(int(l.sum())) - np.repeat(starts, l)).astype(np.uint32))
        return self


def _chop(seq: np.ndarray, max_node: int):
    return [seq[i:i + max_node] for i in range(0, len(seq), max_node)]


def make_variant_graph(length=1_000_000, n_snp=800, n_ins=100, n_del=100, n_haps=8, seed=2,
                       max_node=32, min_spacing=50, name="variants", repeat_unit=0, repeat_copies=0) -> SynthGraph:
    """Config 1/2 graph: random backbone, SNP/insertion/deletion sites >= min_spacing apart,
    haplotypes pick the alt allele with p = 0.5 per site, nodes chopped to <= max_node bp.
    repeat_unit / repeat_copies: the backbone carries that many exact copies of one random unit (minimizers with
    several hits, several clusters per read, ties for the shuffles) — test graphs only."""
    rng = np.random.default_rng(seed)
    ref = BASES[rng.integers(0, 4, size=length)]
    if repeat_unit and repeat_copies:
        unit = BASES[np.random.default_rng(seed + 1000).integers(0, 4, size=repeat_unit)]
        stride = length // repeat_copies
        for c in range(repeat_copies):
            at = c * stride + stride // 3
            ref[at:at + repeat_unit] = unit
    n_var = n_snp + n_ins + n_del
    # variant positions at least min_spacing apart, away from the ends
    sites = []
    if n_var:
        span = length - 2 * min_spacing
        assert span // n_var > min_spacing + 12, "too many variants for this length"
        cell = span // n_var
        sites = [min_spacing + i * cell + int(rng.integers(0, cell - min_spacing - 11)) for i in range(n_var)]
    kinds = np.array([0] * n_snp + [1] * n_ins + [2] * n_del)
    rng.shuffle(kinds)

    node_seqs, slots, slot_min = [], [], []
    site_slot = []

    def add_node(arr) -> int:
        node_seqs.append(bytes(arr).decode())
        return len(node_seqs)

    def add_backbone(arr):
        for piece in _chop(arr, max_node):
            nid = add_node(piece)
            slots.append([nid]); slot_min.append(len(piece))

    cursor = 0
    for pos, kind in zip(sites, kinds):
        if kind == 0:      # SNP at pos
            add_backbone(ref[cursor:pos])
            alt = BASES[(int(np.searchsorted(BASES, ref[pos])) + int(rng.integers(1, 4))) % 4]
            a0 = add_node(ref[pos:pos + 1]); a1 = add_node(np.array([alt], dtype=np.uint8))
            slots.append([a0, a1]); slot_min.append(1)
            cursor = pos + 1
        elif kind == 1:    # insertion before pos
            add_backbone(ref[cursor:pos])
            ins = BASES[rng.integers(0, 4, size=int(rng.integers(1, 11)))]
            a1 = add_node(ins)
            slots.append([0, a1]); slot_min.append(0)
            cursor = pos
        else:              # deletion of ref[pos:pos+L]
            add_backbone(ref[cursor:pos])
            L = int(rng.integers(1, 11))
            a0 = add_node(ref[pos:pos + L])
            slots.append([a0, 0]); slot_min.append(0)
            cursor = pos + L
        site_slot.append(len(slots) - 1)
    add_backbone(ref[cursor:])

    prefix = np.zeros(len(slots) + 1, dtype=np.int64)
    prefix[1:] = np.cumsum(slot_min)
    dist = np.zeros(len(node_seqs) + 1, dtype=capi.dist_dt)
    dist["allele"] = 0xFFFF
    for s, alleles in enumerate(slots):
        for a, nid in enumerate(alleles):
            if nid == 0:
                continue
            dist[nid]["x_in"] = prefix[s]
            dist[nid]["x_out"] = prefix[s + 1]
            dist[nid]["slot"] = s
            dist[nid]["allele"] = 0xFFFF if len(alleles) == 1 else a
    # haplotypes
    site_set = set(site_slot)
    paths = []
    for h in range(n_haps):
        choice = rng.integers(0, 2, size=len(slots))
        p = []
        for s, alleles in enumerate(slots):
            nid = alleles[0] if s not in site_set else alleles[int(choice[s])]
            if nid:
                p.append(2 * nid)
        paths.append(p)
    return SynthGraph(node_seqs, paths, dist, slots=slots, name=name).finish()


def make_tiny_graph(seed=1) -> SynthGraph:
    """Config 1: ~1 kbp, 2 SNPs + one 3-bp deletion, 2 haplotypes."""
    g = make_variant_graph(length=1000, n_snp=2, n_ins=0, n_del=1, n_haps=2, seed=seed, name="tiny")
    return g


def make_branchy_graph(n_layers=31250, n_haps=16, seed=4, name="branchy") -> SynthGraph:
    """Config 4: every 8 bp one base has four alleles (four 8-bp nodes per layer); haplotypes
    choose uniformly, so successive layers are densely connected."""
    rng = np.random.default_rng(seed)
    node_seqs, slots = [], []
    dist = np.zeros(4 * n_layers + 1, dtype=capi.dist_dt)
    base = BASES[rng.integers(0, 4, size=(n_layers, 8))]
    var_pos = rng.integers(0, 8, size=n_layers)
    for layer in range(n_layers):
        alleles = []
        for a in range(4):
            s = base[layer].copy(); s[var_pos[layer]] = BASES[a]
            node_seqs.append(bytes(s).decode())
            nid = len(node_seqs)
            alleles.append(nid)
            dist[nid] = (8 * layer, 8 * (layer + 1), layer, a, 0)
        slots.append(alleles)
    choice = rng.integers(0, 4, size=(n_haps, n_layers))
    paths = [[2 * slots[l][int(choice[h, l])] for l in range(n_layers)] for h in range(n_haps)]
    return SynthGraph(node_seqs, paths, dist, slots=slots, name=name).finish()


@dataclass
class ReadSet:
    reads: np.ndarray       # uint8 [n, L]  (ASCII)
    quals: np.ndarray       # uint8 [n, L]  (Phred, no offset)
    hap: np.ndarray         # truth: haplotype
    pos: np.ndarray         # truth: start on the haplotype's forward strand (of the sampled window)
    rev: np.ndarray         # truth: read is the reverse complement of the haplotype window
    paired: bool = False

    @property
    def n(self):
        return self.reads.shape[0]

    @property
    def length(self):
        return self.reads.shape[1]


def _apply_substitutions(rng, reads, rate):
    if rate <= 0:
        return reads
    mask = rng.random(reads.shape) < rate
    idx = np.searchsorted(BASES, reads[mask])
    reads[mask] = BASES[(idx + rng.integers(1, 4, size=idx.shape)) % 4]
    return reads


def _apply_indels(rng, read: np.ndarray, ins_rate, del_rate, L):
    out = []
    for c in read:
        r = rng.random()
        if r < del_rate:
            continue
        out.append(c)
        if r > 1.0 - ins_rate:
            out.append(BASES[rng.integers(0, 4)])
    out = np.array(out[:L], dtype=np.uint8)
    if len(out) < L:
        out = np.concatenate([out, BASES[rng.integers(0, 4, size=L - len(out))]])
    return out


def simulate_reads(g: SynthGraph, n: int, length=150, sub_rate=0.01, ins_rate=0.0, del_rate=0.0,
                   seed=11, qual=30) -> ReadSet:
    """Single-end reads sampled uniformly from the haplotypes, both strands."""
    rng = np.random.default_rng(seed)
    hap = rng.integers(0, len(g.paths), size=n)
    extra = 16 if (ins_rate > 0 or del_rate > 0) else 0
    reads = np.empty((n, length), dtype=np.uint8)
    pos = np.empty(n, dtype=np.int64)
    for h in range(len(g.paths)):
        sel = np.nonzero(hap == h)[0]
        if len(sel) == 0:
            continue
        hs = g.hap_seq[h]
        p = rng.integers(0, len(hs) - length - extra, size=len(sel))
        pos[sel] = p
        win = hs[p[:, None] + np.arange(length + extra)[None, :]]
        if extra:
            for j, i in enumerate(sel):
                reads[i] = _apply_indels(rng, win[j], ins_rate, del_rate, length)
        else:
            reads[sel] = win
    rev = rng.random(n) < 0.5
    reads[rev] = revcomp_bytes(reads[rev])
    reads = _apply_substitutions(rng, reads, sub_rate)
    quals = np.full((n, length), qual, dtype=np.uint8)
    return ReadSet(reads, quals, hap, pos, rev)


def simulate_pairs(g: SynthGraph, n_pairs: int, length=150, frag_mean=400.0, frag_sd=50.0,
                   sub_rate=0.002, seed=22, qual=30, indel_rate=0.0) -> ReadSet:
    """Paired-end reads, inward orientation: reads[2i] is mate 1, reads[2i+1] mate 2.
    indel_rate (per base): at most one 1-bp insertion or deletion per read, half each, applied in
    haplotype coordinates of the read's window (a deletion skips a base, an insertion adds a random one)."""
    rng = np.random.default_rng(seed)
    hap = rng.integers(0, len(g.paths), size=n_pairs)
    frag = np.clip(np.rint(rng.normal(frag_mean, frag_sd, size=n_pairs)).astype(np.int64), length, None)
    reads = np.empty((2 * n_pairs, length), dtype=np.uint8)
    pos = np.empty(2 * n_pairs, dtype=np.int64)
    flip = rng.random(n_pairs) < 0.5       # fragment from the reverse strand
    for h in range(len(g.paths)):
        sel = np.nonzero(hap == h)[0]
        if len(sel) == 0:
            continue
        hs = g.hap_seq[h]
        start = rng.integers(0, len(hs) - frag[sel].max() - 1, size=len(sel))
        rstart = start + frag[sel] - length
        ar = np.arange(length)[None, :]
        if indel_rate > 0:
            wins = []
            for w0 in (start, rstart):
                hit = rng.random(len(sel)) < 1.0 - (1.0 - indel_rate) ** length
                kind = np.where(hit, rng.integers(1, 3, size=len(sel)), 0)
                p = rng.integers(1, length - 1, size=len(sel))
                ins = BASES[rng.integers(0, 4, size=len(sel))]
                off = ((kind[:, None] == 1) & (ar >= p[:, None])).astype(np.int64) - ((kind[:, None] == 2) & (ar > p[:, None])).astype(np.int64)
                win = hs[w0[:, None] + ar + off]
                win = np.where((kind[:, None] == 2) & (ar == p[:, None]), ins[:, None], win)
                wins.append(win)
            left, right = wins[0], revcomp_bytes(wins[1])
        else:
            left = hs[start[:, None] + ar]
            right = revcomp_bytes(hs[rstart[:, None] + ar])
        f = flip[sel]
        m1 = np.where(f[:, None], right, left)
        m2 = np.where(f[:, None], left, right)
        reads[2 * sel] = m1
        reads[2 * sel + 1] = m2
        pos[2 * sel] = np.where(f, rstart, start)
        pos[2 * sel + 1] = np.where(f, start, rstart)
    rev = np.empty(2 * n_pairs, dtype=bool)
    rev[0::2] = flip
    rev[1::2] = ~flip
    reads = _apply_substitutions(rng, reads, sub_rate)
    quals = np.full(reads.shape, qual, dtype=np.uint8)
    return ReadSet(reads, quals, np.repeat(hap, 2), pos, rev, paired=True)


def truth_seeds(g: SynthGraph, rs: ReadSet, read_offsets=(0, 37, 74, 111), false_seeds=0, seed=5):
    """(node, diag) seeds taken from the true placement of each read (for stage-level tests of
    the extension kernel, independent of the minimizer stage).  Reverse-strand reads get
    seeds on reverse-oriented nodes."""
    rng = np.random.default_rng(seed)
    L = rs.length
    node_len = np.array([0] + [len(s) for s in g.node_seqs], dtype=np.int64)
    items = []
    for i in range(rs.n):
        h, p = int(rs.hap[i]), int(rs.pos[i])
        sd = []
        for ro in read_offsets:
            if ro >= L:
                continue
            if not rs.rev[i]:
                nid, off = int(g.hap_node[h][p + ro]), int(g.hap_off[h][p + ro])
                sd.append((2 * nid, ro - off))
            else:
                hb = p + L - 1 - ro          # haplotype base under read offset ro
                nid, off = int(g.hap_node[h][hb]), int(g.hap_off[h][hb])
                sd.append((2 * nid + 1, ro - (int(node_len[nid]) - 1 - off)))
        for _ in range(false_seeds):
            nid = int(rng.integers(1, len(g.node_seqs) + 1))
            off = int(rng.integers(0, node_len[nid]))
            sd.append((2 * nid + int(rng.integers(0, 2)), int(rng.integers(0, L)) - off))
        items.append((i, sd))
    return items


def make_nested_graph(n_items=60, n_haps=10, seed=7, max_depth=2, name="nested") -> SynthGraph:
    """A chain whose sites are everything the flat chain-of-bubbles model could not hold: alleles of several nodes,
    bubbles nested inside alleles, deletions (empty alleles), sites that touch each other without a backbone node
    between them.  No payload is passed: the index builder derives chains, sites and site tables from the graph the
    haplotypes span (gb_index_build with dist = NULL)."""
    rng = np.random.default_rng(seed)
    node_seqs = []

    def new_node(lo=1, hi=32):
        node_seqs.append(bytes(BASES[rng.integers(0, 4, size=int(rng.integers(lo, hi + 1)))]).decode())
        return len(node_seqs)

    def gen_chain(n, depth, backbone_ends):
        items = []
        for i in range(n):
            r = rng.random()
            if (backbone_ends and (i == 0 or i == n - 1)) or r < 0.45 or depth == 0 and r < 0.6:
                items.append(("node", new_node(8, 32)))
            elif r < 0.65:
                items.append(("site", [[("node", new_node(1, 1))] for _ in range(int(rng.integers(2, 5)))]))          # SNP
            elif r < 0.78:
                items.append(("site", [[("node", new_node(1, 12))], []]))                                               # indel
            elif r < 0.9 or depth == 0:
                items.append(("site", [[("node", new_node(1, 32)) for _ in range(int(rng.integers(1, 4)))] for _ in range(int(rng.integers(2, 4)))]))   # multi-node alleles
            else:
                items.append(("site", [gen_chain(int(rng.integers(1, 4)), depth - 1, False) for _ in range(int(rng.integers(2, 4)))]))                    # nested
        return items

    chain = gen_chain(n_items, max_depth, True)

    def walk(items, choose, out):
        for it in items:
            if it[0] == "node":
                out.append(2 * it[1])
            else:
                alleles = it[1]
                walk(alleles[choose(id(it), len(alleles))], choose, out)

    paths = []
    for h in range(n_haps):
        memo = {}

        def choose(key, n, h=h, memo=memo):
            if key not in memo:
                memo[key] = (h + len(memo)) % n if h < 4 else int(rng.integers(0, n))
            return memo[key]
        p = []
        walk(chain, choose, p)
        paths.append(p)
    used = sorted({v >> 1 for p in paths for v in p})
    remap = {old: new + 1 for new, old in enumerate(used)}           # drop nodes no haplotype visits
    node_seqs = [node_seqs[o - 1] for o in used]
    paths = [[2 * remap[v >> 1] for v in p] for p in paths]
    return SynthGraph(node_seqs, paths, None, slots=None, name=name).finish()
