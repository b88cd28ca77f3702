set -x
cd $GRAFT_REPO_ROOT
python tools/sanitize_batch.py
bash tools/sanitize_gpu.sh gpurun_out/sanitizer
