# Round 2: GPU parity suite with the periodic record / Analyze / sweep / NCCL statistics tests
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_gpu2.log 2>&1; tail -15 gpurun_out/r2_pytest_gpu2.log
