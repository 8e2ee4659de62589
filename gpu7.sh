cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:fast_ -s 7 -c 7 -o gpurun_out/prof_all python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/ncu_all.err
ls -la gpurun_out | tail -3
