# ROUND-1 RECORD: the batch that produced profiles/r01_*; some of the variants it switches between
# (QPWC_CORR_VARIANT=rowpair/packed, tools/ablate_rp.py) were removed in round 2.  Current batch: tools/final_capture_r02c.sh
set -x
cd $GRAFT_REPO_ROOT
python bench.py --steps 500 --warmup 20 > gpurun_out/bench_n1_r01.json 2> gpurun_out/bench_n1_r01.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01.json 2> gpurun_out/bench_ref_r01.err
python tools/level_bench.py --bwd --iters 20 --json gpurun_out/levels_r01.json > gpurun_out/levels_r01.txt 2>&1
python tools/ab_corr.py > gpurun_out/ab_corr_r01.txt 2>&1
QPWC_CORR_VARIANT=rowpair python tools/ablate_rp.py > gpurun_out/ablate_rowpair_r01.txt 2>&1
# launch list (cold-cache, serialised): shares only
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
# full captures of the dominant kernels at the finest level
ncu --set full --import-source on --clock-control none -k regex:corr_fwd_tiled -s 2 -c 1 -f -o gpurun_out/prof_corr_l4_r01 python tools/prof_one.py --op corr --level 4 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:corr_fwd_tiled -s 2 -c 1 -f -o gpurun_out/prof_fused_l4_r01 python tools/prof_one.py --op fused --level 4 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:warp_fwd -s 2 -c 1 -f -o gpurun_out/prof_warp_l4_r01 python tools/prof_one.py --op warp --level 4 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:corr_bwd_tiled -c 1 -f -o gpurun_out/prof_corr_bwd_l4_r01 python tools/prof_one.py --op corr_bwd --level 4 --iters 1 > /dev/null 2>&1
QPWC_CORR_VARIANT=rowpair ncu --set full --import-source on --clock-control none -k regex:rowpair -s 2 -c 1 -f -o gpurun_out/prof_rowpair_l4_r01 python tools/prof_one.py --op corr --level 4 > /dev/null 2>&1
QPWC_CORR_VARIANT=packed ncu --set full --import-source on --clock-control none -k regex:corr_fwd_tiled -s 2 -c 1 -f -o gpurun_out/prof_corr_packed_l4_r01 python tools/prof_one.py --op corr --level 4 > /dev/null 2>&1
ls -la gpurun_out | tail -20
