mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" python bench.py --no-e2e --no-cpu --no-parity > gpurun_out/ab_$tag.log 2> gpurun_out/ab_$tag.err; echo "$tag rc=$?"; }
run hint0 EOD_TMA_WAIT_HINT_NS=0
run hint1000 EOD_TMA_WAIT_HINT_NS=1000
run hint20000 EOD_TMA_WAIT_HINT_NS=20000
run hint0b EOD_TMA_WAIT_HINT_NS=0
run hint1000b EOD_TMA_WAIT_HINT_NS=1000
