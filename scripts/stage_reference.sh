#!/bin/bash
# Stage the UNMODIFIED reference under the git-ignored baseline/_ref/ so that it travels to the GPU box with a gpurun
# snapshot.  Nothing of it is committed; tests/test_dropin_l2.py skips when it is absent.
set -e
cd "$(dirname "$0")/.."
mkdir -p baseline/_ref
cp -r /root/reference/mmdet /root/reference/configs baseline/_ref/
find baseline/_ref -name "__pycache__" -type d -prune -exec rm -rf {} +
du -sh baseline/_ref
