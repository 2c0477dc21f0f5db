cd $GRAFT_REPO_ROOT
python bench.py --steps 300 --warmup 10 --no-train --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_fj.json; python -c "
import json;d=json.load(open('gpurun_out/bench_fj.json'));print(d['ms_per_step'],d['config']['serial_levels_ms_per_step'],d['roofline']['avg_launch_ms'],d['clocks']);print(d['e2e'])"
timeout 300 python -m pytest tests -m gpu -x -q -k "host" 2>&1 | tail -2
