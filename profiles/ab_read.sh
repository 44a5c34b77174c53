mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "read or smoke or bench_conf or stress" 2>&1 | tail -3
python profiles/prof_read.py > gpurun_out/prof_read_v4.json 2>gpurun_out/prof_read_v4.err; cat gpurun_out/prof_read_v4.json
PROF_E=16 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:read_pool_kernel -c 6 --csv --log-file gpurun_out/read_v4_inst.csv python profiles/prof_read.py > /dev/null 2>&1; tail -8 gpurun_out/read_v4_inst.csv | cut -c1-400
