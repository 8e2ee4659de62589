#!/bin/bash
# Round-1 evidence run on ONE B200 (under gpurun): parity suite, bench (own arm
# and reference arm), ncu launch list and one `--set full` capture of a whole
# convolution.  The .ncu-rep stays in /tmp (it exceeds the 64 MiB gpurun_out
# limit); its raw/details pages are exported as CSV text instead.
#   gpurun --timeout 1500 -- 'bash profiles/collect_r01.sh'
# (SKIP_REF=1 / SKIP_NCU=1 leave out the reference arm / the ncu captures)
# then:  python profiles/summarize_ncu.py gpurun_out/ncu_r01_raw.csv
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01_n1.json 2> gpurun_out/bench_r01_n1.err
tail -c 300 gpurun_out/bench_r01_n1.json
if [ -z "$SKIP_REF" ]; then
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_reference.json 2> gpurun_out/bench_r01_reference.err
tail -c 300 gpurun_out/bench_r01_reference.json
fi
python profiles/run_configs.py > gpurun_out/configs_r01.jsonl 2> gpurun_out/configs_r01.err
cat gpurun_out/configs_r01.jsonl | cut -c1-200
[ -n "$SKIP_NCU" ] && exit 0
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu --set full --clock-control none -k regex:fast_ -s 7 -c 7 -o /tmp/prof_r01_all \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu -i /tmp/prof_r01_all.ncu-rep --page raw --csv > gpurun_out/ncu_r01_raw.csv 2>/dev/null
ncu -i /tmp/prof_r01_all.ncu-rep --page details --csv > gpurun_out/ncu_r01_details.csv 2>/dev/null
ls -la gpurun_out | tail -12
