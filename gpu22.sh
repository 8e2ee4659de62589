cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_conv.py -x -q -m gpu -k "stride or multnone" 2>&1 | tail -15
