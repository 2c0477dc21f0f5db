# Round-2 evidence batch (run under gpurun from the repo root): bench records, per-level timings, engine
# A/B, micro-benchmark, config-4 sweep, ncu launch list and full captures of the dominant kernels.
set -x
cd $GRAFT_REPO_ROOT
T="timeout 400"
$T python bench.py --steps 500 --warmup 20 > gpurun_out/bench_n1_r02.json 2> gpurun_out/bench_n1_r02.err
$T python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02.json 2> gpurun_out/bench_ref_r02.err
$T python tools/level_bench.py --bwd --iters 20 --json gpurun_out/levels_r02.json > gpurun_out/levels_r02.txt 2>&1
$T python tools/ab_engines.py > gpurun_out/engines_r02.txt 2>&1
timeout 60 ./tools/ubench/tf32x3_tile > gpurun_out/tf32x3_ubench_r02.txt 2>&1
$T python tools/sweep_cfg4.py gpurun_out/cfg4_sweep_r02.json > gpurun_out/cfg4_sweep_r02.txt 2>&1
$T python tools/ab_warp_bwd.py > gpurun_out/warp_bwd_variants_r02.txt 2>&1
$T python tools/train_hotpath_profile.py 64 > gpurun_out/train_hotpath_r02.txt 2>&1
$T python tools/tc_trace.py 32 > gpurun_out/tc_trace_r02.txt 2>&1
# launch list of the timed steps (cold-cache, serialised: compare shares)
$T ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 5 --warmup 3 --no-graph --no-cpu-baseline --no-train > gpurun_out/ncu_launch_r02.log 2>&1
# full captures at the finest level / a streaming level
$T ncu --set full --import-source on --clock-control none -k regex:corr_fwd_tc_res -s 2 -c 1 -f -o gpurun_out/prof_tc_res_l4_r02 python tools/prof_one.py --op corr --level 4 > /dev/null 2>&1
$T ncu --set full --import-source on --clock-control none -k regex:corr_fwd_tc_stream -s 2 -c 1 -f -o gpurun_out/prof_tc_stream_l3_r02 python tools/prof_one.py --op corr --level 3 > /dev/null 2>&1
$T ncu --set full --import-source on --clock-control none -k regex:warp_fwd -s 2 -c 1 -f -o gpurun_out/prof_warp_l4_r02 python tools/prof_one.py --op warp --level 4 > /dev/null 2>&1
$T ncu --set full --import-source on --clock-control none -k regex:warp_bwd -s 1 -c 1 -f -o gpurun_out/prof_warp_bwd_l4_r02 python tools/prof_one.py --op warp_bwd --level 4 --iters 2 > /dev/null 2>&1
$T ncu --set full --import-source on --clock-control none -k regex:corr_bwd_tiled -c 1 -f -o gpurun_out/prof_corr_bwd_l4_r02 python tools/prof_one.py --op corr_bwd --level 4 --iters 1 > /dev/null 2>&1
ls -la gpurun_out | tail -30
