set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r1_v4n.json 2> gpurun_out/bench_r1_v4n.err; tail -c 600 gpurun_out/bench_r1_v4n.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_ref2.json 2> gpurun_out/bench_r1_ref2.err; tail -c 300 gpurun_out/bench_r1_ref2.json
timeout 600 python bench.py --workload C4 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r1_c4_v4n.json 2> gpurun_out/bench_r1_c4_v4n.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_c4_v4n.json')); print('C4', d['kernel_ms'], d['roofline']['frac'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_v4n.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1; tail -2 gpurun_out/ncu_l.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_v4 -s 2 -c 1 -f -o gpurun_out/prof_fused_v4n python profiles/prof_driver.py fused 3 > gpurun_out/ncu_f.log 2>&1; tail -2 gpurun_out/ncu_f.log
python profiles/batch12.py 2>&1 | tail -3
