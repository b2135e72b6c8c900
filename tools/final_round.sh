python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --profile --steps 1 > gpurun_out/profile_plain_r1z.json 2> gpurun_out/profile_plain_r1z.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r1z.csv python bench.py --profile --steps 1 > gpurun_out/ncu_launches_r1z.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_final_n1.json') if l.startswith('{')][0]
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'][:40], d['roofline']['frac'])"
