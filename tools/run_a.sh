cd $GRAFT_REPO_ROOT
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python tools/ab_engines.py 2>&1 | tail -7
python bench.py --steps 300 --warmup 10 --no-train --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_fj.json; python -c "
import json;d=json.load(open('gpurun_out/bench_fj.json'));print(d['ms_per_step'],d['config']['serial_levels_ms_per_step'],d['config']['upflow_path'],d['roofline']['avg_launch_ms'],d['clocks'])"
