"""Transcribe the fragment-length fixture of the reference into tests/golden/fragment_lengths.json.
Reads /root/reference/src/unittest/minimizer_mapper.cpp:37-108 ("Fragment length distribution gets
reasonable value": a vector of observed pair distances with a heavy outlier tail, registered when
<= max_fragment_length (2000), with REQUIRE(std_dev() <= 400)).  Development container only; the tests
and the GPU box never read the reference tree.
usage: python scripts/extract_fragment_vectors.py"""
import json, re
from pathlib import Path

SRC = Path("/root/reference/src/unittest/minimizer_mapper.cpp")
OUT = Path(__file__).resolve().parents[1] / "tests" / "golden" / "fragment_lengths.json"
lines = SRC.read_text().splitlines()
start = next(i for i, l in enumerate(lines) if "vector<int64_t> distances" in l)
end = next(i for i in range(start, len(lines)) if "};" in lines[i])
body = " ".join(lines[start:end + 1])
body = body[body.index("{") + 1: body.rindex("}")]
distances = [int(x) for x in re.findall(r"-?\d+", body)]
doc = {"source": f"unittest/minimizer_mapper.cpp:{start + 1}-{end + 1}", "max_fragment_length": 2000,
       "distribution": {"maximum_sample_size": 1000, "reestimation_frequency": 1000, "robust_estimation_fraction": 0.95,
                        "source": "minimizer_mapper.cpp:72"},
       "require": {"std_dev_at_most": 400}, "distances": distances}
OUT.write_text(json.dumps(doc) + "\n")
print(len(distances), "distances ->", OUT)
