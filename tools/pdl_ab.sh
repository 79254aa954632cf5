#!/bin/bash
# A/B of programmatic dependent launch (B200SEG_PDL, csrc/common.cuh) on one B200: parity tests with the switch on,
# then the default bench and the batch-1 / batch-8 inference sweep with the switch off and on.
#   gpurun --timeout 420 -- 'bash tools/pdl_ab.sh'
mkdir -p gpurun_out/pdl
export PYTHONUNBUFFERED=1
B200SEG_PDL=1 timeout 240 python -m pytest tests/test_gpu_engine.py tests/test_gpu_config0.py tests/test_gpu_determinism.py \
    -x -q -m gpu > gpurun_out/pdl/tests_pdl1.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/pdl/tests_pdl1.log
for m in 0 1; do
  B200SEG_PDL=$m timeout 150 python bench.py --no-cpu-baseline --steps 10 --warmup 3 \
      > gpurun_out/pdl/bench_pdl$m.json 2> gpurun_out/pdl/bench_pdl$m.err
  echo "bench pdl=$m rc=$?"
  B200SEG_PDL=$m timeout 90 python tools/infer_sweep.py --batches 1,8 --sides 256 --reps 20 \
      > gpurun_out/pdl/infer_pdl$m.jsonl 2> gpurun_out/pdl/infer_pdl$m.err
  echo "infer pdl=$m rc=$?"
done
tail -3 gpurun_out/pdl/tests_pdl1.log
python - <<'P'
import json
for m in (0, 1):
    try:
        d = json.loads(open(f"gpurun_out/pdl/bench_pdl{m}.json").read().strip().splitlines()[-1])
        print("pdl", m, "img/s", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), d["clocks"]["sm_mhz"])
    except Exception as e:
        print("pdl", m, "bench unreadable", e)
    try:
        for ln in open(f"gpurun_out/pdl/infer_pdl{m}.jsonl"):
            d = json.loads(ln); print("pdl", m, "infer b", d["batch"], {k: v for k, v in d.items() if "ms" in k or "img" in k})
    except Exception as e:
        print("pdl", m, "infer unreadable", e)
P
