timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_regressions.py -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu --no-ref-gpu --no-extra --no-e2e > /tmp/b.json 2>/dev/null; python - <<PY
import json
d = json.loads(open("/tmp/b.json").read().strip().splitlines()[-1])
print("main(pairs,2)", round(d["value"]), "solves/s", round(d["ms_per_step"], 4), "ms", {k: round(v["us_per_launch"], 1) for k, v in d["kernels"].items()})
PY
tools/gpu_ab.sh ab libgato_b200_pmb3.so libgato_b200_nopairs.so
