set -x
nvidia-smi -L
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_conv.py -x -q -m gpu 2>&1 | tail -30
python -m pytest tests/test_gpu_pad.py -x -q -m gpu 2>&1 | tail -30
