#!/bin/bash
# A/B of programmatic dependent launch (B200SEG_PDL, csrc/common.cuh) on one B200: parity tests with the switch set to
# $1 (default 2 = size rule), then small-batch training, the inference sweep and the default bench with it off and on.
#   gpurun --timeout 420 -- 'bash tools/pdl_ab.sh 2'
M=${1:-2}
mkdir -p gpurun_out/pdl
export PYTHONUNBUFFERED=1
B200SEG_PDL=$M timeout 240 python -m pytest tests/test_gpu_engine.py tests/test_gpu_train_loop.py tests/test_gpu_determinism.py \
    -x -q -m gpu > gpurun_out/pdl/tests_pdl$M.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/pdl/tests_pdl$M.log
for m in 0 $M; do
  B200SEG_PDL=$m timeout 150 python bench.py --model R2AttU_Net --t 2 --batch 4 --no-cpu-baseline --steps 20 --warmup 5 \
      > gpurun_out/pdl/bench_r2attu_b4_pdl$m.json 2> gpurun_out/pdl/bench_r2attu_b4_pdl$m.err
  echo "bench r2attu b4 pdl=$m rc=$?"
  B200SEG_PDL=$m timeout 90 python tools/infer_sweep.py --batches 1,8,64 --sides 256 --reps 20 \
      > gpurun_out/pdl/infer_pdl$m.jsonl 2> gpurun_out/pdl/infer_pdl$m.err
  echo "infer pdl=$m rc=$?"
done
B200SEG_PDL=$M timeout 150 python bench.py --no-cpu-baseline --steps 10 --warmup 3 \
    > gpurun_out/pdl/bench_pdl$M.json 2> gpurun_out/pdl/bench_pdl$M.err
echo "bench pdl=$M rc=$?"
tail -3 gpurun_out/pdl/tests_pdl$M.log
M=$M python - <<'P'
import json, os
M = os.environ["M"]
def last(path):
    return json.loads(open(path).read().strip().splitlines()[-1])
for m in ("0", M):
    try:
        d = last(f"gpurun_out/pdl/bench_r2attu_b4_pdl{m}.json")
        print("pdl", m, "R2AttU b4 img/s", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), d["clocks"]["sm_mhz"])
    except Exception as e:
        print("pdl", m, "bench unreadable", e)
    try:
        for ln in open(f"gpurun_out/pdl/infer_pdl{m}.jsonl"):
            d = json.loads(ln); print("pdl", m, "infer b", d["batch"], d["ms_per_batch"])
    except Exception as e:
        print("pdl", m, "infer unreadable", e)
try:
    d = last(f"gpurun_out/pdl/bench_pdl{M}.json")
    print("pdl", M, "AttU b64 img/s", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), d["clocks"]["sm_mhz"])
except Exception as e:
    print("bench unreadable", e)
P
