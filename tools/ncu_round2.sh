#!/bin/bash
# Round-2 ncu captures (run under gpurun on ONE GPU).  Outputs under gpurun_out/; summaries are made here afterwards
# with tools/ncu_summary.py.  Numbers printed by a run under ncu are never bench values.
set -x
O=gpurun_out
NCU="ncu --clock-control none"
# every launch of the default bench command (device times: compare shares, not absolutes)
timeout 600 $NCU --metrics gpu__time_duration.sum -k regex:"la3d|mask_scan|sample_kernel|fit_kernel|fit_all|rle_decode|lift_|prep_kernel|peer_sync" \
    -c 600 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_bench_under_ncu.log 2>&1
# one full capture per kernel of a step, of the lift and of the all-pixels fit
timeout 600 $NCU --set full --import-source on -k regex:"mask_scan_kernel|sample_kernel|fit_kernel" -s 6 -c 3 -o $O/r2_step2048 -f python tools/profile_r2.py step > $O/r2_ncu_step.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:"mask_scan_kernel|sample_kernel|fit_kernel" -s 6 -c 3 -o $O/r2_step256 -f python tools/profile_r2.py step256 >> $O/r2_ncu_step.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:"lift_bulk_kernel" -s 2 -c 1 -o $O/r2_lift_cfg2 -f python tools/profile_r2.py lift > $O/r2_ncu_lift.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:"lift_bulk_kernel" -s 5 -c 1 -o $O/r2_lift_cfg4 -f python tools/profile_r2.py lift >> $O/r2_ncu_lift.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:"fit_all_kernel" -s 1 -c 5 -o $O/r2_fitall -f python tools/profile_r2.py fitall > $O/r2_ncu_fitall.log 2>&1
ls -la $O/*.ncu-rep | tail
