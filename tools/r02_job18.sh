#!/bin/bash
# final build on an 8-GPU box: every multi-GPU correctness test (NCCL dist at 2/4/8, one-process multi entry point, C harness, CLI)
mkdir -p gpurun_out/r02; cd /root/repo; O=gpurun_out/r02
timeout 500 python -m pytest tests/test_gpu_dist.py tests/test_gpu_parity.py tests/test_host_cli.py -m gpu -q -k "dist or multi or harness or cli or shards or gpus" > $O/j18_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/j18_pytest.log
