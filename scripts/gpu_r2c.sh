# round 2: per-launch device times of one LMM step in real conditions (CUDA events, warm caches), 16- and 8-element geometries
set -x
for e in 16; do
  rm -f gpurun_out/launch_dump_e$e.txt
  FMC_OPTIONS=tape_elems=$e,profile=1 FMC_PROFILE_DUMP=gpurun_out/launch_dump_e$e.txt timeout -s KILL 300 python - <<PY
import sys, os
sys.path.insert(0, "finmath-lib-cuda-extensions_b200"); sys.path.insert(0, ".")
import finmath_cuda as fc
from finmath_cuda import _capi as capi
from finmath_cuda.workloads import DriverLib
fc.ensure_init()
m = DriverLib().lmm(1 << 20, 80, 0.5, 1, 31415, 0, (0, 1 << 20))
for _ in range(3): m.step()
capi.profile_read()
m.step()
print(capi.profile_read())
PY
done
