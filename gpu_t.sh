cd $GRAFT_REPO_ROOT
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=1',d['value'],'conv/s',d['ms_per_step'],'ms', {(k['pass'],k['op']):round(k['ms_per_step'],2) for k in d['kernels']})"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fast_conv_rows -s 1 -c 1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "dram__|gpu__time" 
cat > /tmp/nb.py <<'PY'
import os
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
import torch, torch.distributed as dist
torch.cuda.set_device(0)
dist.init_process_group("nccl", device_id=torch.device("cuda",0))
t=torch.ones(4,device="cuda"); dist.all_reduce(t); torch.cuda.synchronize()
print("STDOUT-LINE")
dist.destroy_process_group()
PY
echo "NCCL_DEBUG env: [$NCCL_DEBUG]"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29631 /tmp/nb.py 2>/dev/null | head -5
