#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list of the bench command and
# one `ncu --set full` capture of the main kernels of one step at 500x500x80.  From the repo root:
#   gpurun --timeout 2400 -- 'bash profiles/run_gpu.sh <tag> [tests] [bench] [launches] [full]'
# Only text comes back (gpurun_out/ is capped at 64 MiB): the .ncu-rep is exported to CSV on the box.
tag=${1:-rX}; shift
what="${*:-tests bench launches full}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
MAIN='^(k_tend_|k_advect_scalar_fast|k_advance_|k_ssp_3d|k_prep_3d|k_ssf|k_pd_|k_calc_p_rho|k_calc_ww|k_calc_coef_w|k_update_scalar|k_zero3|k_invert_cqw|k_mom_|k_ph_fast)'
for w in $what; do
  case $w in
    tests)
      timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
      tail -3 gpurun_out/${tag}_pytest.log ;;
    bench)
      timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
      tail -c 3500 gpurun_out/${tag}_bench.json ;;
    sanitize)
      # memcheck of one whole step of every parity case through the C-ABI (slow: minutes)
      timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k full_step \
        > gpurun_out/${tag}_sanitize.log 2>&1
      echo "sanitize rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${tag}_sanitize.log | tail -3 ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
        python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-second-workload > gpurun_out/${tag}_launches.log 2>&1
      tail -1 gpurun_out/${tag}_launches.log | cut -c1-200 ;;
    full)
      # skip start_domain and the warm-up step, then one whole step of the main kernels
      timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$MAIN" --launch-skip ${SKIP:-70} --launch-count ${COUNT:-70} \
        -o /tmp/${tag}_full -f python bench.py --steps 1 --warmup 1 --nx 500 --ny 500 --no-cpu-baseline --no-second-workload > gpurun_out/${tag}_full.log 2>&1
      tail -1 gpurun_out/${tag}_full.log | cut -c1-200
      ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>/dev/null
      for kn in ${SRC_KERNELS:-k_mom_u_fast k_advance_uv k_advance_w k_advance_mu_t k_advect_scalar_fast k_pd_tend_fast}; do
        ncu -i /tmp/${tag}_full.ncu-rep --page source --csv -k regex:"^$kn" --launch-count 1 > gpurun_out/${tag}_src_${kn}.csv 2>/dev/null
      done
      sz=$(du -sm /tmp/${tag}_full.ncu-rep | cut -f1)
      if [ "$sz" -lt 40 ]; then cp /tmp/${tag}_full.ncu-rep gpurun_out/; fi
      du -sm gpurun_out ;;
  esac
done
