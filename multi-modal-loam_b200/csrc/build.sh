#!/bin/bash
# Build libmmloam_b200.so for sm_100a (nvcc cross-compiles without a GPU).
# -fmad=false: float32/float64 expressions must round exactly like the reference's
# baseline x86-64 build (no FMA contraction) for bit-exact feature labels and k-NN sets.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="${MML_OUT:-$HERE/../libmmloam_b200.so}"   # MML_OUT / MML_OBJ: variant builds for A/B runs (scratch/)
OBJ="${MML_OBJ:-$HERE/_obj}"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -Xcompiler -O3 ${MML_EXTRA_NVCC_FLAGS}"
mkdir -p "$OBJ"
pids=()
for f in extract geometry splitvoxel framesort associate accumulate odometry localmap globalmap window windowsolve msgpack capi; do
  if [ ! -f "$OBJ/$f.o" ] || [ -n "$(find "$HERE" -maxdepth 1 \( -name '*.cu' -o -name '*.cuh' \) -newer "$OBJ/$f.o" 2>/dev/null)" ] || [ "$HERE/../../include/mmloam_b200.h" -nt "$OBJ/$f.o" ]; then
    # accumulate.cu is pure float64 normal-equation arithmetic checked to 1e-9 relative (no bit-exact float32
    # thresholds inside): it may contract multiply-adds into DFMA
    if [ "$f" = "accumulate" ] || [ "$f" = "windowsolve" ]; then FF="${FLAGS/-fmad=false/-fmad=true}"; else FF="$FLAGS"; fi
    # window.cu carries the host side of the sliding-window solve (IMU factors under forward-mode differentiation,
    # dense (15 W)-dim dogleg): let the host compiler vectorise it (every B200 host CPU has AVX2 + FMA)
    if [ "$f" = "window" ]; then FF="$FF -Xcompiler -mavx2 -Xcompiler -mfma -Xcompiler -ffp-contract=fast"; fi
    $NVCC $FF -c "$HERE/$f.cu" -o "$OBJ/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "$OBJ"/{extract,geometry,splitvoxel,framesort,associate,accumulate,odometry,localmap,globalmap,window,windowsolve,msgpack,capi}.o -lcudart
echo "built $OUT"
