#!/bin/sh
# TEST / BENCH INFRASTRUCTURE.  Compiles the reference's own MSDeformAttn CUDA kernel (header-only) from where it lies under
# the reference tree into oracle/_ref/libmsda_ref.so.  Only possible where the reference tree exists (build container);
# the binary is git-ignored and travels to the GPU box with the snapshot.  Usage: sh oracle/build_ref.sh [reference root]
set -e
REF="${1:-${TAIR_REF:-/root/reference}}"
HERE="$(cd "$(dirname "$0")" && pwd)"
HDR="$REF/testr/adet/layers/csrc/DeformAttn"
[ -f "$HDR/ms_deform_im2col_cuda.cuh" ] || { echo "reference kernel header not found under $REF" >&2; exit 2; }
TORCH_INC="$(python -c 'import torch, os; print(os.path.join(os.path.dirname(torch.__file__), "include"))')"
mkdir -p "$HERE/_ref"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC -cudart static \
     -DCUDA_HAS_FP16=1 -D__CUDA_NO_HALF_OPERATORS__ -D__CUDA_NO_HALF_CONVERSIONS__ -D__CUDA_NO_HALF2_OPERATORS__ \
     -I "$HDR" -I "$TORCH_INC" -I "$TORCH_INC/torch/csrc/api/include" \
     "$HERE/msda_ref_launcher.cu" -o "$HERE/_ref/libmsda_ref.so"
echo "built $HERE/_ref/libmsda_ref.so"
