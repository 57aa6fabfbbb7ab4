"""include/giraffe_b200.hpp (the C++ mirror of vg's MinimizerMapper / FragmentLengthDistribution seams over the C ABI)
compiles with -Wall -Werror against the library, and its host-side behaviour — the distribution mirror, the parameter
aliases, exceptions instead of a fallback when there is no device — is what tests/cpp/shim_check.cpp expects."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_cpp_mirror_compiles_links_and_behaves(tmp_path):
    from vg_b200 import capi
    capi.load_library()                                            # built (no fallback)
    exe = tmp_path / "shim_check"
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-o", str(exe), str(ROOT / "tests" / "cpp" / "shim_check.cpp"),
           "-L", str(ROOT / "vg_b200"), "-lgiraffe_b200", f"-Wl,-rpath,{ROOT / 'vg_b200'}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0 and "shim ok" in run.stdout, run.stdout + run.stderr


@pytest.mark.gpu
def test_cpp_mirror_maps_like_the_c_abi(tmp_path):
    """giraffe_b200::MinimizerMapper::map / map_batch / map_paired_batch in a C++ program vs Device.map_arrays here."""
    import numpy as np
    import helpers as H
    from vg_b200 import capi, synth
    g = synth.make_variant_graph(length=30000, n_snp=48, n_ins=6, n_del=6, n_haps=4, seed=3)
    index = g.build_index()
    index.save(tmp_path / "g.gbflat")
    rs = synth.simulate_pairs(g, 40, sub_rate=0.01, seed=7)
    with open(tmp_path / "reads.txt", "w") as f:
        for i in range(rs.n):
            f.write(f"r{i} {bytes(rs.reads[i]).decode()}\n")
    exe = tmp_path / "shim_gpu_check"
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-o", str(exe), str(ROOT / "tests" / "cpp" / "shim_gpu_check.cpp"),
           "-L", str(ROOT / "vg_b200"), "-lgiraffe_b200", f"-Wl,-rpath,{ROOT / 'vg_b200'}"]
    assert subprocess.run(cmd, capture_output=True, text=True).returncode == 0
    run = subprocess.run([str(exe), str(tmp_path / "g.gbflat"), str(tmp_path / "reads.txt")], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    lines = [l.split() for l in run.stdout.splitlines()]
    dev = capi.Device(index)
    se = H.gpu_map(dev, rs.reads, rs.quals)
    pe = H.gpu_map(dev, rs.reads, rs.quals, _paired_defaults(), paired=True)

    def expect(res, i):
        a = res[0][i]; mapped = a["flags"] & 1
        first = res[1][int(a["mapping_off"])] if mapped else None
        return [str(int(a["score"])), str(int(a["mapq"])), str(int(first["node"])) if mapped else "-1", str(int(first["offset"])) if mapped else "0", str(int(a["n_mappings"]))]

    se_lines = [l for l in lines if l[0] == "SE"]; pe_lines = [l for l in lines if l[0] == "PE"]; one = [l for l in lines if l[0] == "ONE"]
    assert len(se_lines) == rs.n and len(pe_lines) == rs.n and len(one) == 1
    for i in range(rs.n):
        assert se_lines[i][1] == f"r{i}" and se_lines[i][2:] == expect(se, i)
        assert pe_lines[i][1] == f"r{i}" and pe_lines[i][2:] == expect(pe, i)
    assert one[0][2:] == expect(se, 0)
    dev.close()


def _paired_defaults():
    import helpers as H
    p = H.default_map_params()          # vg defaults: rescue on; the C++ mirror uses gb_map_params_default too
    p.fragment_mean = 400.0; p.fragment_stdev = 50.0
    return p
