#!/bin/bash
# scratch/build_variant.sh NAME [SRC_DIR] ["EXTRA NVCC FLAGS"]: builds scratch/v/NAME.so from SRC_DIR (default: the tree's csrc)
set -euo pipefail
NAME=$1; SRC=${2:-$(dirname "$0")/../pytracer_b200/csrc}; FLAGS=${3:-}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
RT_OUT=$ROOT/scratch/v/$NAME.so RT_BUILD_DIR=/tmp/rt_build_$NAME RT_EXTRA_FLAGS="$FLAGS" bash $SRC/build.sh
