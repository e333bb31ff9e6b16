#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Compiles the reference's own deformable-convolution op -- UNMODIFIED, from the two source files where
# they lie under /root/reference (mmdet/ops/dcn/src/deform_conv_cuda.cpp + deform_conv_cuda_kernel.cu; the reference builds
# them through setup.py:101 for compute_70) -- for sm_100a into oracle/_ref/deform_conv_cuda.so (git-ignored, travels to the
# GPU box).  It is the checker of the UPSNetFPN-subnet row: tests/test_dcn.py and
# tests/golden/make_golden_dcn.py import it on the B200; nothing under slotvps_b200/ ever does.
# The only addition is an include path with a one-line THC/THCAtomics.cuh shim (that header left PyTorch).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${SLOTVPS_REFERENCE:-/root/reference}"
SRC="$REF/mmdet/ops/dcn/src"
[ -f "$SRC/deform_conv_cuda.cpp" ] || { echo "reference sources not found under $REF (GPU box: use the prebuilt oracle/_ref)"; exit 0; }
OUT="$HERE/_ref/deform_conv_cuda.so"
mkdir -p "$HERE/_ref"
if [ -f "$OUT" ] && [ "$OUT" -nt "$SRC/deform_conv_cuda.cpp" ] && [ "$OUT" -nt "$SRC/deform_conv_cuda_kernel.cu" ] && [ "${1:-}" != "--force" ]; then
  echo "up to date: $OUT"; exit 0
fi
TORCH="$(python -c 'import torch, os; print(os.path.dirname(torch.__file__))')"
PYINC="$(python -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
"${NVCC:-/usr/local/cuda/bin/nvcc}" -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -Xcompiler -fPIC -shared -w \
  -DTORCH_EXTENSION_NAME=deform_conv_cuda -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=1 \
  -I"$HERE/shim" -I"$TORCH/include" -I"$TORCH/include/torch/csrc/api/include" -I"$PYINC" \
  "$SRC/deform_conv_cuda.cpp" "$SRC/deform_conv_cuda_kernel.cu" \
  -L"$TORCH/lib" -ltorch -ltorch_cpu -ltorch_cuda -lc10 -lc10_cuda -ltorch_python -o "$OUT"
echo "built $OUT"
