"""The headline numbers quoted in DESIGN.md / README.md are the ones in the committed bench lines under profiles/."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.load(f)


def _text(name):
    with open(os.path.join(ROOT, name), encoding="utf-8") as f:
        return re.sub(r"[\s  ]", "", f.read())        # "178 840" == "178840"


def test_design_quotes_the_committed_single_gpu_lines():
    design = _text("DESIGN.md")
    for name in ("r2_bench_c3.json", "r2_bench_c4.json", "r2_bench_c5.json"):
        d = _line(name)
        assert d["data"] == "synthetic" and d["n_gpus"] == 1
        assert str(round(d["value"])) in design, (name, round(d["value"]))
        assert str(round(d["e2e"]["value"])) in design, (name, round(d["e2e"]["value"]))
    c3 = _line("r2_bench_c3.json")
    assert c3["config"]["byte_identical_to_single_gpu"] is True
    assert c3["roofline"]["bound"] == "tensor" and c3["cpu_baseline"]["kind"] in ("reference", "port")


def test_multi_gpu_lines_are_identical_to_the_single_gpu_lists():
    sha = _line("r2_bench_c3.json")["config"]["sha1_lists"]
    design = _text("DESIGN.md")
    for name in ("r2_bench_n2.json", "r2_bench_n4.json", "r2_bench_n8.json", "r2_bench_n8_single_process.json"):
        d = _line(name)
        assert d["config"]["sha1_lists"] == sha == d["config"]["sha1_single_gpu"] == d["config"]["sha1_e2e"], name
        assert d["scaling"] == "strong" and d["n_gpus"] in (2, 4, 8)
    n8 = _line("r2_bench_n8.json")
    assert f"{round(n8['value'] / 1000)}k" in design, round(n8["value"] / 1000)       # "1 394 k"
    assert f"{round(n8['e2e']['value'] / 1000)}k" in design


def test_readme_headline_matches_the_c3_line():
    readme = _text("README.md")
    c3 = _line("r2_bench_c3.json")
    assert f"{c3['value'] / 1000:.1f}k" in readme and f"{c3['e2e']['value'] / 1000:.1f}k" in readme
