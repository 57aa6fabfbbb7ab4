"""The reference's clusterer unit vectors that fit the chain-of-bubbles index model
(src/unittest/snarl_seed_clusterer.cpp:174-532), run against the oracle's definition of a cluster:
components of "unoriented minimum distance <= limit" through the 16-byte distance payload
(snarl_seed_clusterer.hpp:15-50).  The GPU clustering is compared with this oracle by every map parity test."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
from vg_b200 import capi


def _index(node_seqs, slots, paths):
    """slots: per chain position the allele node ids (0 = empty allele); payload as synth.make_variant_graph."""
    lens = [0] + [len(s) for s in node_seqs]
    slot_min = [min(lens[n] if n else 0 for n in alleles) for alleles in slots]
    prefix = np.zeros(len(slots) + 1, dtype=np.int64); prefix[1:] = np.cumsum(slot_min)
    dist = np.zeros(len(node_seqs) + 1, dtype=capi.dist_dt)
    dist["allele"] = 0xFFFF
    comp = 0
    for s, alleles in enumerate(slots):
        for a, nid in enumerate(alleles):
            if nid:
                dist[nid]["x_in"] = prefix[s]; dist[nid]["x_out"] = prefix[s + 1]; dist[nid]["slot"] = s
                dist[nid]["allele"] = 0xFFFF if len(alleles) == 1 else a
                dist[nid]["component"] = comp
    return dist


def _cluster(index, positions, reads, read_limit, fragment_limit=0):
    lib = H.oracle_lib()
    lib.oracle_cluster_positions.argtypes = [C.POINTER(capi.FlatIndex), C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                             C.c_void_p, C.c_void_p]
    lib.oracle_cluster_positions.restype = None
    node = np.array([2 * i + int(rev) for i, rev, _ in positions], dtype=np.uint32)
    off = np.array([o for _, _, o in positions], dtype=np.uint32)
    rd = np.array(reads, dtype=np.uint32)
    rl = np.zeros(len(node), dtype=np.uint32); fl = np.zeros(len(node), dtype=np.uint32)
    lib.oracle_cluster_positions(C.byref(index.view), len(node), capi.ptr(node), capi.ptr(off), capi.ptr(rd), read_limit, fragment_limit, capi.ptr(rl), capi.ptr(fl))
    return rl.tolist(), fl.tolist()


# :174-192 / :319-340: GCA -(T|G)- CTGA -(GCA|T)- T, and the same plus a node of its own component
CHAIN_SEQS = ["GCA", "T", "G", "CTGA", "GCA", "T", "T"]
CHAIN_SLOTS = [[1], [2, 3], [4], [5, 6], [7]]
CHAIN_PATHS = [[2, 4, 8, 10, 14], [2, 6, 8, 12, 14]]


@pytest.fixture(scope="module")
def chain():
    return capi.HostIndex(CHAIN_SEQS, CHAIN_PATHS, _index(CHAIN_SEQS, CHAIN_SLOTS, CHAIN_PATHS), k=5, w=3)


@pytest.fixture(scope="module")
def chain_plus_component():
    seqs = CHAIN_SEQS + ["TTTTTTTTT"]
    dist = _index(seqs, CHAIN_SLOTS, CHAIN_PATHS)
    dist[8]["x_in"] = 0; dist[8]["x_out"] = 9; dist[8]["slot"] = 0; dist[8]["component"] = 1
    return capi.HostIndex(seqs, CHAIN_PATHS + [[16]], dist, k=5, w=3)


def test_simple_chain_one_cluster_on_the_same_node(chain):
    rl, _ = _cluster(chain, [(4, False, 0), (4, False, 1), (4, False, 3)], [0, 0, 0], 2)         # :198-214
    assert len(set(rl)) == 1


def test_simple_chain_opposite_sides_of_a_snp(chain):
    pos = [(2, False, 0), (3, False, 0), (5, False, 0)]
    assert len(set(_cluster(chain, pos, [0, 0, 0], 10)[0])) == 1                                  # :215-233
    assert len(set(_cluster(chain, pos, [0, 0, 0], 4)[0])) == 3                                   # :234-250


def test_simple_chain_two_reads(chain):
    rl, fl = _cluster(chain, [(2, False, 0), (3, False, 0), (5, False, 0)], [0, 0, 1], 5, 5)      # :251-283
    assert len({rl[0], rl[1]}) == 2 and fl[0] == fl[2]
    rl, fl = _cluster(chain, [(5, False, 0), (6, False, 0), (1, False, 0)], [0, 0, 1], 10, 10)    # :284-317
    assert len({rl[0], rl[1]}) == 2 and fl[0] == fl[2]


def test_chain_with_a_second_component(chain_plus_component):
    ix = chain_plus_component
    assert len(set(_cluster(ix, [(4, False, 0), (4, False, 1), (4, False, 3), (8, False, 3)], [0] * 4, 2)[0])) == 2      # :343-361
    pos = [(2, False, 0), (3, False, 0), (5, False, 0), (8, False, 0)]
    assert len(set(_cluster(ix, pos, [0] * 4, 10)[0])) == 2                                       # :362-378
    assert len(set(_cluster(ix, pos, [0] * 4, 4)[0])) == 4                                        # :379-395
    rl, fl = _cluster(ix, [(2, False, 0), (3, False, 0), (5, False, 0)], [0, 0, 1], 5, 5)         # :396-428
    assert len({rl[0], rl[1]}) == 2 and fl[0] == fl[2]


def test_long_snarl_in_chain():
    """:465-532: GGC -(GCA|-)- GCA -(16 bp | GCA)- 19 bp; a reverse-strand position on node 2."""
    seqs = ["GGC", "GCA", "GCAGCACATGCACATC", "GCA", "GCAAGCACATGCACATCCA", "GCA"]
    slots = [[1], [6, 0], [2], [3, 4], [5]]
    paths = [[2, 12, 4, 6, 10], [2, 4, 8, 10]]
    index = capi.HostIndex(seqs, paths, _index(seqs, slots, paths), k=5, w=3)
    pos = [(2, True, 0), (3, False, 8), (5, False, 0)]
    assert len(set(_cluster(index, pos, [0, 0, 0], 5)[0])) == 2
    assert len(set(_cluster(index, pos, [0, 0, 0], 2)[0])) == 3
