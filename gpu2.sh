cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -3
nproc; free -g | head -2
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_generic.json 2> gpurun_out/bench_generic.err
cat gpurun_out/bench_generic.json; tail -5 gpurun_out/bench_generic.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --chunk 4 > gpurun_out/bench_generic_c4.json 2>> gpurun_out/bench_generic.err
cat gpurun_out/bench_generic_c4.json
