mkdir -p gpurun_out
timeout 300 build/umma_probe 50000 200 2048 i8 > gpurun_out/probe9.txt 2>&1; echo "probe exit $?" >> gpurun_out/probe9.txt
grep -E "m=50000|layout tests|full tests|accuracy" gpurun_out/probe9.txt | tail -8
timeout 900 python -m pytest tests/test_zz_lowrank.py -m gpu -q --maxfail=10 -rfEs --tb=short > gpurun_out/pytest_o.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_o.txt
tail -3 gpurun_out/pytest_o.txt
timeout 600 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg5_o.json 2> gpurun_out/bench_cfg5_o.err; echo "exit $?" >> gpurun_out/bench_cfg5_o.err
python -c "
import json; j=json.load(open('gpurun_out/bench_cfg5_o.json')); print('cfg5', j['value'], j['ms_per_step'], j['setup_ms'], j['setup_cold_ms'], j['setup'], j['e2e']['value'])"
