#!/bin/bash
# Copies the reference's four hot-path scripts into gpurun_in/reference_scripts/ -- a git-ignored directory that gpurun
# ships with the snapshot -- so that tests/test_gpu_reference_scripts.py can run them UNMODIFIED on a B200.  They are
# job inputs, never committed.   usage: tools/stage_reference_scripts.sh [/root/reference]
set -e
src=${1:-/root/reference}
dst="$(dirname "$0")/../gpurun_in/reference_scripts"
mkdir -p "$dst/util"
cp "$src/util/qdrant_manager.py" "$dst/util/"
cp "$src/31_clip_embedding_and_save_vector.py" "$src/32_create_delegate_vector.py" "$src/33_run_all_experiments.py" "$dst/"
ls -l "$dst" "$dst/util"
