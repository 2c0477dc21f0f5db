# Round-2 (final batch) evidence: bench records, per-level timings, engine A/B, config-4 sweep, training hot
# path, kernel time line, ncu launch list and full captures of the two tensor-core kernels.
set -x
cd $GRAFT_REPO_ROOT
T="timeout 400"
$T python -m pytest tests -m gpu -x -q > gpurun_out/r02c_gpu_tests.log 2>&1
$T python bench.py --steps 500 --warmup 20 > gpurun_out/bench_n1_r02c.json 2> gpurun_out/bench_n1_r02c.err
$T python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02c.json 2> gpurun_out/bench_ref_r02c.err
$T python tools/level_bench.py --bwd --iters 20 --json gpurun_out/levels_r02c.json > gpurun_out/levels_r02c.txt 2>&1
$T python tools/ab_engines.py > gpurun_out/engines_r02c.txt 2>&1
$T python tools/sweep_cfg4.py gpurun_out/cfg4_sweep_r02c.json > gpurun_out/cfg4_sweep_r02c.txt 2>&1
$T python tools/tc_d8.py > gpurun_out/tc_d8_r02c.txt 2>&1
$T python tools/warp_nchw_bench.py > gpurun_out/warp_nchw_r02c.log 2>&1
$T python tools/train_hotpath_profile.py 64 > gpurun_out/train_hotpath_r02c.txt 2>&1
$T python tools/tc_trace.py 32 > gpurun_out/tc_trace_r02c.txt 2>&1
$T python tools/ablate_tc.py > gpurun_out/tc_ablation_r02c.txt 2>&1
# launch list of the timed steps (cold-cache, serialised: compare shares)
$T ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02c.csv python bench.py --steps 5 --warmup 3 --no-graph --no-cpu-baseline --no-train > gpurun_out/ncu_launch_r02c.log 2>&1
$T ncu --set full --import-source on --clock-control none -k regex:corr_fwd_tc_res -s 2 -c 1 -f -o gpurun_out/prof_tc_res_l4_r02c python tools/prof_one.py --op corr --level 4 > /dev/null 2>&1
$T ncu --set full --import-source on --clock-control none -k regex:corr_fwd_tc_stream -s 2 -c 1 -f -o gpurun_out/prof_tc_stream_l3_r02c python tools/prof_one.py --op corr --level 3 > /dev/null 2>&1
tail -3 gpurun_out/r02c_gpu_tests.log
ls -la gpurun_out | tail -20
