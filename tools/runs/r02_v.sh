# round 2, GPU call V: ncu full-set and launch list on the FINAL build
mkdir -p gpurun_out
timeout 300 python tools/profile_step.py --K 3 > gpurun_out/v_profile_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'level_fwd|level_bwd|contract_f16' -s 16 -c 16 -f -o gpurun_out/r02_v_full python tools/profile_step.py --K 3 > gpurun_out/v_ncu.log 2>&1
tail -2 gpurun_out/v_ncu.log
timeout 300 python tools/profile_step.py --K 25 --steps 1 > gpurun_out/v_profile_plain25.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_v_launches.csv python tools/profile_step.py --K 25 --steps 1 > gpurun_out/v_ncu_launches.log 2>&1
tail -2 gpurun_out/v_ncu_launches.log; wc -l gpurun_out/r02_v_launches.csv
