set -x
O=gpurun_out; mkdir -p $O
timeout 120 python profiles/prof_fuse.py > $O/prof_fuse.json 2>$O/prof_fuse.err; tail -1 $O/prof_fuse.json
PROF_ONLY=tc PROF_ROUNDS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:project_fuse_persistent -s 2 -c 2 -f -o $O/prof_fuse_v2 python profiles/prof_fuse.py > $O/ncu_fuse.log 2>&1; echo rc=$?
tail -3 $O/ncu_fuse.log
