#!/usr/bin/env bash
# compute-sanitizer passes over the protocol-heavy kernels: run on a GPU box with
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# Logs land in gpurun_out/sanitize_<tool>.log; copy the summaries you want judged into profiles/.
# racecheck   shared-memory hazards between the producer / MMA / epilogue warps (generic-proxy accesses; TMA and
#             tcgen05 traffic goes through the async proxy, which racecheck does not model — the mbarrier protocol is
#             covered by synccheck and by the result checks inside tools/sanitize_cases.py)
# synccheck   invalid bar.sync / mbarrier usage, divergent barriers
# memcheck    out-of-bounds / misaligned global and shared accesses, leaked device memory
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rc_all=0
for tool in memcheck synccheck racecheck; do
  log=gpurun_out/sanitize_${tool}.log
  echo "== compute-sanitizer --tool ${tool}" | tee "${log}"
  timeout 600 compute-sanitizer --tool "${tool}" --error-exitcode 86 --print-limit 20 \
      python tools/sanitize_cases.py all >> "${log}" 2>&1
  rc=$?
  echo "exit code ${rc}" | tee -a "${log}"
  tail -4 "${log}"
  [ "${rc}" -ne 0 ] && rc_all=1
done
exit ${rc_all}
