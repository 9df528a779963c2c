#!/bin/bash
# Round-2 measurement pass on one B200 (run under gpurun): the whole -m gpu suite, then every bench line.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2_gputest_full.log 2>&1; tail -4 gpurun_out/r2_gputest_full.log | cut -c1-300
run() { name=$1; shift; python bench.py "$@" > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.err || tail -3 gpurun_out/r2_bench_$name.err; python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/r2_bench_{n}.json"))
    r = d.get("roofline", {})
    print(n, "value", round(d["value"], 1), d["unit"], "e2e", round(d["e2e"]["value"], 1), "frac", r.get("frac"), "achieved", r.get("achieved"), r.get("unit"),
          "knn_share", r.get("knn_share_of_step"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "clk", (d.get("clocks") or {}).get("sm_mhz"))
except Exception as e:
    print(n, "FAILED", e)
PY
}
run c3 --steps 10 --warmup 3
run ref_c3 --impl reference --steps 3 --warmup 1
run orb_tensor --workload orb --steps 10 --warmup 3
run orb_popc --workload orb --orb-engine popc --steps 5 --warmup 3 --no-cpu-baseline
run knnmatch --workload knnmatch --steps 3 --warmup 1
run c4 --workload c4 --steps 10 --warmup 3 --no-cpu-baseline
run extract --workload extract --steps 5 --warmup 3
run c5 --workload c5 --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 3
run extract_orb --workload extract --detector ORB --steps 5 --warmup 3
