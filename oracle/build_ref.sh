#!/usr/bin/env bash
# TEST INFRASTRUCTURE — not part of the product.
#
# Builds the UNMODIFIED reference (LumiOwO/RT-Octree) straight from the sources where they lie
# under /root/reference into oracle/_ref/ (git-ignored, but it travels to the GPU box).  No
# reference source is copied into the tracked tree; the reference's own CMake is not used (it
# hard-requires OpenGL/GLFW/GLEW: renderer/CMakeLists.txt:123-125,233-261).
#
# Outputs (all under oracle/_ref/):
#   volrend_headless   the reference CLI (renderer/main_headless.cpp) for sm_100a
#   ref_driver         oracle/ref_driver.cpp (ours) linked against the same reference objects; it calls
#                      volrend::launch_renderer / Denoiser::denoise and dumps aux + final float image
#   volrend.ptx        PTX of the reference render kernel (read to pin the op sequence)
#   _denoiser_ref.so   the reference's Python extension (denoiser/extension/bindings.cpp + filtering.cu) as the torch
#                      extension module `_denoiser_ref`: the unmodified `filtering_autograd` forward + backward
#   libref_cpu.so      host-compiled reference trace_ray/query/SH functions (oracle/ref_cpu_shim.cpp)
#   ref_pose_dump      the reference's own _recenter_poses (llff) made callable (oracle/ref_pose_shim.cpp)
#
# Two build-time tweaks only (both outside the forward path): `-include cstdint` for imwrite.cpp
# (gcc 13) and `.type()` -> `.scalar_type()` in the training-backward dispatch of filtering.cu.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF_ROOT:-/root/reference}"
OUT="$HERE/_ref"
WHAT="${1:-all}"
if [ ! -d "$REF/renderer" ]; then
  echo "build_ref.sh: $REF not present (GPU box?) - using prebuilt oracle/_ref if any"; exit 0
fi
mkdir -p "$OUT/obj" "$OUT/inc/volrend" "$OUT/patched"
R="$REF/renderer"
TP="$(python3 -c 'import torch,os;print(os.path.dirname(torch.__file__))')"
CUDA_HOME="${CUDA_HOME:-/usr/local/cuda}"
ARCH="-gencode arch=compute_100a,code=sm_100a"

sed -e 's/@VOLREND_VERSION_MAJOR@/0/;s/@VOLREND_VERSION_MINOR@/0/;s/@VOLREND_VERSION_PATCH@/1/' \
    -e 's/@_VOLREND_CUDA_@//;s/@_VOLREND_PNG_@/\/\/ /' "$R/common.hpp.in" > "$OUT/inc/volrend/common.hpp"
INC="-I $OUT/inc -I $R/include -I $R/3rdparty -I $R/3rdparty/cnpy -I $R/3rdparty/glm -I $R/3rdparty/misc -I $CUDA_HOME/include"
TI="-I $TP/include -I $TP/include/torch/csrc/api/include"

newer() { [ ! -e "$2" ] || [ "$1" -nt "$2" ]; }

build_cuda_ref() {
  for f in volrend n3tree common; do
    if newer "$R/src/cuda/$f.cu" "$OUT/obj/${f}_cu.o"; then
      nvcc -std=c++17 -O3 $ARCH -lineinfo $INC -c "$R/src/cuda/$f.cu" -o "$OUT/obj/${f}_cu.o" &
    fi
  done
  if newer "$R/src/cuda/volrend.cu" "$OUT/volrend.ptx"; then
    nvcc -std=c++17 -O3 -arch=compute_100a -ptx $INC "$R/src/cuda/volrend.cu" -o "$OUT/volrend.ptx" &
  fi
  if newer "$REF/denoiser/extension/filtering.cu" "$OUT/obj/filtering.o"; then
    sed 's/grad_guidance\.type()/grad_guidance.scalar_type()/' "$REF/denoiser/extension/filtering.cu" > "$OUT/patched/filtering.cu"
    cp "$REF/denoiser/extension/filtering.h" "$OUT/patched/"
    nvcc -std=c++17 -O3 $ARCH $TI -I "$OUT/patched" -c "$OUT/patched/filtering.cu" -o "$OUT/obj/filtering.o" &
  fi
  for f in src/n3tree src/camera src/opts 3rdparty/cnpy/cnpy main_headless; do
    o="$OUT/obj/$(basename $f).o"
    if newer "$R/$f.cpp" "$o"; then g++ -std=c++17 -O2 -w $INC -c "$R/$f.cpp" -o "$o" & fi
  done
  if newer "$R/src/imwrite.cpp" "$OUT/obj/imwrite.o"; then
    g++ -std=c++17 -O2 -w -include cstdint $INC -c "$R/src/imwrite.cpp" -o "$OUT/obj/imwrite.o" &
  fi
  if newer "$R/src/denoiser/denoiser.cpp" "$OUT/obj/denoiser.o"; then
    g++ -std=c++17 -O2 -w $INC $TI -I "$OUT/patched" -c "$R/src/denoiser/denoiser.cpp" -o "$OUT/obj/denoiser.o" &
  fi
  g++ -std=c++17 -O2 -w $INC $TI -I "$OUT/patched" -c "$HERE/ref_driver.cpp" -o "$OUT/obj/ref_driver.o" &
  wait
  LIBOBJ="$OUT/obj/n3tree.o $OUT/obj/camera.o $OUT/obj/opts.o $OUT/obj/imwrite.o $OUT/obj/cnpy.o \
          $OUT/obj/n3tree_cu.o $OUT/obj/common_cu.o $OUT/obj/volrend_cu.o $OUT/obj/denoiser.o $OUT/obj/filtering.o"
  LINK="-L$TP/lib -ltorch -ltorch_cpu -ltorch_cuda -lc10 -lc10_cuda -L$CUDA_HOME/lib64 -lcudart -lz -lpthread -ldl -Wl,-rpath,$TP/lib"
  g++ -o "$OUT/volrend_headless" "$OUT/obj/main_headless.o" $LIBOBJ $LINK
  g++ -o "$OUT/ref_driver" "$OUT/obj/ref_driver.o" $LIBOBJ $LINK
  echo "built $OUT/volrend_headless and $OUT/ref_driver"
  # the reference's own llff pose recentring (anonymous-namespace functions of main_headless.cpp), callable: ref_pose_shim.cpp
  g++ -std=c++17 -O2 -w $INC -I "$R" -c "$HERE/ref_pose_shim.cpp" -o "$OUT/obj/ref_pose_shim.o"
  g++ -o "$OUT/ref_pose_dump" "$OUT/obj/ref_pose_shim.o" $LIBOBJ $LINK
  echo "built $OUT/ref_pose_dump"
  # the reference's own torch extension (what denoiser/network.py:12-46 JIT-builds), under a non-clashing module name
  PYINC="$(python3 -c 'import sysconfig;print(sysconfig.get_paths()["include"])')"
  if newer "$OUT/patched/filtering.cu" "$OUT/_denoiser_ref.so"; then
    nvcc -std=c++17 -O3 $ARCH $TI -I "$PYINC" -I "$OUT/patched" -Xcompiler -fPIC -DTORCH_EXTENSION_NAME=_denoiser_ref \
         -D_GLIBCXX_USE_CXX11_ABI=1 -c "$OUT/patched/filtering.cu" -o "$OUT/obj/filtering_pic.o" &
    g++ -std=c++17 -O2 -w -fPIC $TI -I "$CUDA_HOME/include" -I "$PYINC" -I "$OUT/patched" -DTORCH_EXTENSION_NAME=_denoiser_ref \
         -c "$REF/denoiser/extension/bindings.cpp" -o "$OUT/obj/bindings.o" &
    wait
    g++ -shared -o "$OUT/_denoiser_ref.so" "$OUT/obj/bindings.o" "$OUT/obj/filtering_pic.o" $LINK -ltorch_python
    echo "built $OUT/_denoiser_ref.so"
  fi
}

build_cpu_ref() {
  # Host compile of the reference's own device headers (rt_core.cuh etc.) through a macro shim.
  g++ -std=c++17 -O3 -mfma -ffp-contract=fast -fopenmp -fPIC -shared -w -I "$HERE/shim" $INC \
      "$HERE/ref_cpu_shim.cpp" -o "$OUT/libref_cpu.so"
  echo "built $OUT/libref_cpu.so"
}

case "$WHAT" in
  cuda) build_cuda_ref ;;
  cpu)  build_cpu_ref ;;
  all)  build_cpu_ref; build_cuda_ref ;;
esac
