"""Transcribe the WFAExtender unit vectors of the reference into tests/golden/wfa.json.
Reads /root/reference/src/unittest/gbwt_extender.cpp (development container only; the GPU box and the
tests never read the reference tree) and parses the regular structure of its SECTIONs:
sequence, from / to positions, connect | prefix | suffix, optional ErrorModel, and the expectations
check_score(matches, mismatches, gaps, gap_length[, full_length_ends]) / REQUIRE_FALSE(result) /
check_unlocalized_insertion / check_alignment(..., &from | nullptr, &to | nullptr).
usage: python scripts/extract_wfa_vectors.py"""
import json, re, sys
from pathlib import Path

SRC = Path("/root/reference/src/unittest/gbwt_extender.cpp")
OUT = Path(__file__).resolve().parents[1] / "tests" / "golden" / "wfa.json"
text = SRC.read_text().splitlines()

GRAPHS = {
    "linear": {"nodes": {1: "CGC", 2: "GATTACA", 3: "GATTA", 4: "TAT"}, "paths": [[1, 2, 3, 4]]},
    "general": {"nodes": {1: "CGC", 2: "GATTACA", 3: "G", 4: "C", 5: "ATTA", 6: "TG", 7: "TC", 8: "GAA", 9: "CAT", 10: "GTA", 11: "TAT"},
                "paths": [[1, 2, 3, 5, 6, 8, 9, 11], [1, 2, 4, 5, 7, 8, 10, 11]]},
    "cycle": {"nodes": {1: "CGC", 2: "GA", 3: "TAT"}, "paths": [[1, 2, 3], [1, 2, 2, 3], [1, 2, 2, 2, 3]]},
    "multi_cycle": {"nodes": {1: "CATTAG", 2: "GA", 3: "TTA", 4: "CA", 5: "TATAGAGA"},
                    "paths": [[1, 5], [1] + [2, 3, 4] * 10 + [5], [1] + [2, 3, 4] * 11 + [5]]},
}
GRAPH_OF = {"wfa_linear_gbwt": "linear", "wfa_general_gbwt": "general", "wfa_cycle_gbwt": "cycle",
            "wfa_non_diverging_multi_node_cycle_gbwt": "multi_cycle"}

def pos(m):
    return [int(m.group(1)), m.group(2) == "true", int(m.group(3))]

POS = r"\(\s*(\d+)\s*,\s*(true|false)\s*,\s*(\d+)\s*\)"
cases, skipped = [], []
i = 0
# find WFA test cases (the ones tagged [wfa_extender])
while i < len(text):
    line = text[i]
    if line.startswith("TEST_CASE(") and "[wfa_extender]" in line:
        tc_name = re.search(r'TEST_CASE\("([^"]+)"', line).group(1)
        # body until the next TEST_CASE at column 0
        j = i + 1
        while j < len(text) and not text[j].startswith("TEST_CASE("):
            j += 1
        body = text[i:j]
        graph = next((GRAPH_OF[k] for k in GRAPH_OF if any(k + "()" in b for b in body)), None)
        # prelude: statements before the first SECTION
        k = 0
        while k < len(body) and "SECTION(" not in body[k]:
            k += 1
        prelude = body[:k]
        # split sections
        secs, cur = [], None
        depth = 0
        for ln_no, b in enumerate(body[k:], start=i + k + 1):
            if "SECTION(" in b and cur is None:
                cur = {"name": re.search(r'SECTION\("([^"]+)"', b).group(1), "line": ln_no, "lines": []}
                depth = b.count("{") - b.count("}")
                continue
            if cur is not None:
                depth += b.count("{") - b.count("}")
                if depth <= 0:
                    secs.append(cur); cur = None
                else:
                    cur["lines"].append(b)
        for sec in secs:
            src = prelude + sec["lines"]
            joined = "\n".join(src)
            seq = None
            ss = ""
            uses_ss = "std::stringstream ss" in joined
            # sequence
            m = re.search(r'std::string sequence\("([^"]*)"\)', "\n".join(sec["lines"]))
            if m:
                seq = m.group(1)
            elif re.search(r"std::string sequence;", "\n".join(sec["lines"])):
                seq = ""
            if uses_ss:
                # replay `ss << "..."` statements and simple counted loops in order
                li = 0
                while li < len(src):
                    s = src[li]
                    lm = re.search(r"for \(size_t i = 0; i < (\d+); i\+\+\)", s)
                    if lm:
                        n = int(lm.group(1))
                        inner = re.search(r'ss << "([^"]*)"', src[li + 1])
                        if inner is None:
                            seq = None; break
                        ss += inner.group(1) * n
                        li += 2
                        continue
                    for sm in re.finditer(r'ss << "([^"]*)"', s):
                        ss += sm.group(1)
                    li += 1
                if "sequence = ss.str()" in joined:
                    seq = ss
            frm = to = None
            for s in src:
                m = re.search(r"pos_t from" + POS, s)
                if m: frm = pos(m)
                m = re.search(r"pos_t to" + POS, s)
                if m: to = pos(m)
            call = None
            for s in sec["lines"]:
                if "extender.connect(" in s: call = "connect"
                elif "extender.prefix(" in s: call = "prefix"
                elif "extender.suffix(" in s: call = "suffix"
            em = None
            m = re.search(r"ErrorModel errors\s*\{(.*?)\};", joined, re.S)
            if m:
                parts = []
                for ev in re.finditer(r"\{\s*([\d.]+)\s*,\s*(\d+)\s*,\s*(\d+)\s*\}|default_(\w+)\(\)", m.group(1)):
                    if ev.group(4):
                        parts += {"mismatches": [0.03, 1, 6], "gaps": [0.05, 1, 10], "gap_length": [0.1, 1, 20], "distance": [0.1, 10, 200]}[ev.group(4)]
                    else:
                        parts += [float(ev.group(1)), int(ev.group(2)), int(ev.group(3))]
                em = parts if len(parts) == 12 else "unparsed"
            if "ErrorModel model = WFAExtender::default_error_model" in joined:
                # default model with individual fields overwritten (model.gap_length.max = 15; ...)
                em = [0.03, 1, 6, 0.05, 1, 10, 0.1, 1, 20, 0.1, 10, 200]
                ev_i = {"mismatches": 0, "gaps": 1, "gap_length": 2, "distance": 3}; f_i = {"per_base": 0, "min": 1, "max": 2}
                for am in re.finditer(r"model\.(\w+)\.(\w+) = ([\d.]+);", joined):
                    em[3 * ev_i[am.group(1)] + f_i[am.group(2)]] = float(am.group(3)) if am.group(2) == "per_base" else int(am.group(3))
            expect = {}
            for s in sec["lines"]:
                m = re.search(r"check_score\(result, aligner, (.*)\);", s)
                if m:
                    args = [a.strip() for a in m.group(1).split(",")]
                    expect["score_terms"] = args
                if re.search(r"REQUIRE_FALSE\(result\)|REQUIRE\(!\(bool\)\(result\)\)|REQUIRE\(!result\)", s):
                    expect["fail"] = True
                if "check_unlocalized_insertion(" in s:
                    expect["unlocalized_insertion"] = True
                m = re.search(r"check_alignment\(result, sequence, graph, aligner, (&from|nullptr), (&to|nullptr)\)", s)
                if m:
                    expect["check_alignment"] = [m.group(1) == "&from", m.group(2) == "&to"]
                m = re.search(r"REQUIRE\(result\.score == (.*)\);", s)
                if m:
                    expect["score_expr"] = m.group(1)
            ok = graph and seq is not None and call and em != "unparsed" and expect and \
                (call != "connect" or (frm and to)) and (call != "prefix" or to) and (call != "suffix" or frm)
            rec = {"test_case": tc_name, "section": sec["name"], "line": sec["line"]}
            if not ok:
                skipped.append(rec); continue
            if "score_terms" in expect:
                L = len(seq)
                vals = [int(eval(a.replace("sequence.length()", str(L)).replace("sequence.size()", str(L)))) for a in expect["score_terms"]]
                expect["score_terms"] = vals + [0] * (5 - len(vals))
            if "score_expr" in expect:
                expect["score_expr"] = int(eval(expect["score_expr"]))
            rec.update({"graph": graph, "call": call, "sequence": seq, "from": frm if call != "prefix" else None,
                        "to": to if call != "suffix" else None, "error_model": em, "expect": expect})
            cases.append(rec)
        i = j
    else:
        i += 1

doc = {"source": "src/unittest/gbwt_extender.cpp TEST_CASEs tagged [wfa_extender], transcribed by scripts/extract_wfa_vectors.py; "
                 "positions are [node id, is_reverse, offset]; score_terms = [matches, mismatches, gaps, total gap length, full-length ends] "
                 "(check_score, :1390); Aligner defaults match 1, mismatch 4, gap_open 6, gap_extend 1, full_length_bonus 5",
       "graphs": {k: {"nodes": {str(a): b for a, b in v["nodes"].items()}, "paths": v["paths"]} for k, v in GRAPHS.items()},
       "cases": cases, "not_transcribed": skipped}
OUT.write_text(json.dumps(doc, indent=1))
print(len(cases), "cases;", len(skipped), "sections not transcribed:", [(s["section"], s["line"]) for s in skipped][:12])
