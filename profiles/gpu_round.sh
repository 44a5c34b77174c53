#!/bin/bash
# One gpurun call: GPU parity tests, smoke, default bench, ncu launch list, ncu --set full of the two top kernels.
# Usage: gpurun --timeout 1500 -- 'bash profiles/gpu_round.sh <tag>'
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_$TAG.log
python __graft_entry__.py --smoke > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_$TAG.log
# adversarial / randomised sweeps against the oracle (bit-exact kernels) and kernel-vs-kernel (tcgen05 variants)
python profiles/prof_online.py > $O/online_$TAG.json 2>$O/online_$TAG.err; echo "online rc=$?"; tail -1 $O/online_$TAG.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_$TAG.log 2>$O/bench_ref_$TAG.err; echo "bench ref rc=$?"
python bench.py > $O/bench_$TAG.log 2>$O/bench_$TAG.err; echo "bench rc=$?"; tail -c 600 $O/bench_$TAG.err
cat $O/bench_$TAG.log
if [ "$2" != "noncu" ]; then
PROF_E=16 PROF_T=3 python profiles/prof_step.py > $O/prof_plain_$TAG.log 2>&1 && \
PROF_E=16 PROF_T=3 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_$TAG.csv python profiles/prof_step.py > $O/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
PROF_E=16 PROF_T=3 ncu --set full --clock-control none --import-source on -k regex:'write_mean_chw_tma|read_pool' -s 2 -c 4 -f -o $O/prof_$TAG python profiles/prof_step.py > $O/ncu_full_$TAG.log 2>&1
echo "full rc=$?"
PROF_E=64 PROF_T=2 python profiles/prof_step.py > $O/prof_plain64_$TAG.log 2>&1 && \
PROF_E=64 PROF_T=2 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/traffic64_$TAG.csv python profiles/prof_step.py > $O/ncu_traffic_$TAG.log 2>&1
echo "traffic rc=$?"
# projection + fusion (tcgen05): timing without a profiler, then one --set full capture of the persistent kernel
python profiles/prof_fuse.py > $O/prof_fuse_$TAG.json 2>$O/prof_fuse_$TAG.err; echo "fuse rc=$?"; tail -1 $O/prof_fuse_$TAG.json
PROF_ONLY=tc PROF_ROUNDS=2 ncu --set full --clock-control none --import-source on -k regex:project_fuse_persistent -s 2 -c 1 -f -o $O/prof_fuse_$TAG python profiles/prof_fuse.py > $O/ncu_fuse_$TAG.log 2>&1
echo "fuse full rc=$?"
# the other BASELINE configurations and the object-regime stage breakdown (no profiler)
python profiles/bench_configs.py > $O/configs_$TAG.log 2>$O/configs_$TAG.err; echo "configs rc=$?"
python profiles/prof_layouts.py > $O/layouts_$TAG.json 2>$O/layouts_$TAG.err; echo "layouts rc=$?"
python profiles/prof_objects.py > $O/objects_$TAG.json 2>$O/objects_$TAG.err; echo "objects rc=$?"
PROF_E=16 PROF_T=4 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_obj_$TAG.csv python profiles/prof_objects.py > $O/ncu_launches_obj_$TAG.log 2>&1
echo "object launch list rc=$?"
fi
