#!/bin/bash
# Round-2 evidence run on ONE B200 (under gpurun): parity suite, bench (own arm
# and reference arm), the five BASELINE configs, ncu launch list and a
# `--set full` capture of one whole convolution (raw page exported as CSV: the
# .ncu-rep exceeds the 64 MiB gpurun_out limit).
#   gpurun --timeout 1800 -- 'bash profiles/collect_r02.sh'
# then:  python profiles/summarize_ncu.py gpurun_out/ncu_r02_raw.csv r02
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_r02.txt 2>&1; tail -3 gpurun_out/pytest_gpu_r02.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err
tail -c 400 gpurun_out/bench_r02_n1.json; echo
if [ -z "$SKIP_REF" ]; then
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_reference.json 2> gpurun_out/bench_r02_reference.err
tail -c 300 gpurun_out/bench_r02_reference.json; echo
fi
python profiles/run_configs.py > gpurun_out/configs_r02.jsonl 2> gpurun_out/configs_r02.err
cut -c1-160 gpurun_out/configs_r02.jsonl
[ -n "$SKIP_NCU" ] && exit 0
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity > /dev/null 2>&1
ncu --set full --clock-control none -k regex:'fast_|tma_' -s 21 -c 7 -f -o /tmp/prof_r02_all \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > /dev/null 2>&1
ncu -i /tmp/prof_r02_all.ncu-rep --page raw --csv > gpurun_out/ncu_r02_raw.csv 2>/dev/null
cuobjdump -sass fftwpp_b200/lib_fftwpp.so | grep -oE "UTMALDG[.A-Z0-9]*|UTMASTG[.A-Z0-9]*|UBLKCP[.A-Z0-9]*|LDTM[.a-z0-9]*|STTM[.a-z0-9]*|SYNCS[.A-Z0-9]*" | sort | uniq -c > gpurun_out/sass_tma_r02.txt
cat gpurun_out/sass_tma_r02.txt
ls -la gpurun_out | tail -12
