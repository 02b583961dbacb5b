# round 2: ncu of the LMM step (simulation kernel + swaption kernel) for the 16- and 8-element geometries
set -x
for e in 16 8; do
  export FMC_OPTIONS=tape_elems=$e
  timeout -s KILL 300 python benchmarks/profile_lmm.py > gpurun_out/prof_plain_e$e.log 2>&1 && \
  timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:tape_kernel -s 20 -c 2 -o gpurun_out/prof_sim_r2b_e$e -f python benchmarks/profile_lmm.py > gpurun_out/ncu_sim_e$e.log 2>&1
  echo "ncu rc=$?"
  timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:tape_kernel -s 140 -c 2 -o gpurun_out/prof_swp_r2b_e$e -f python benchmarks/profile_lmm.py > gpurun_out/ncu_swp_e$e.log 2>&1
  echo "ncu rc=$?"
done
