#!/bin/bash
# final single-GPU validation of round 2: full GPU suite, bench line, config table, long-row ncu
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r02_final.txt 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_r02_final.txt | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench_r02_final_n1.json 2> gpurun_out/bench_r02_final_n1.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/bench_r02_final_n1.json
timeout 600 python profiles/run_configs.py > gpurun_out/configs_r02_final.jsonl 2> gpurun_out/configs_r02_final.err; echo "configs rc=$?"
cut -c1-330 gpurun_out/configs_r02_final.jsonl
timeout 120 python - > gpurun_out/cfg1_scan.jsonl 2>&1 <<'PY'
import json, sys, torch
sys.path.insert(0, ".")
import fftwpp_b200 as fp
L = 1 << 20
d = [torch.zeros(L, dtype=torch.complex128, device="cuda") for _ in range(2)]
for m in (None, [1024], [2048], [4096]):
    c = fp.HybridConv([L], [2 * L], m=m)
    for _ in range(5):
        c.convolve(d, normalized=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        c.convolve(d, normalized=False)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"cfg1 forced m": m, "params": c.params(0), "ms": e0.elapsed_time(e1) / 50}))
    c.close()
PY
cat gpurun_out/cfg1_scan.jsonl | cut -c1-300
EXP_LONG_ONLY=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fast_conv_rows_long -c 1 -o gpurun_out/ncu_long3 python scripts/exp_long.py > gpurun_out/ncu_long3.log 2>&1; echo "ncu rc=$?"
