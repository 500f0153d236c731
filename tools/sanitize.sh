#!/bin/bash
# Builds AddressSanitizer, UndefinedBehaviorSanitizer and ThreadSanitizer variants of libq3asr.so (host code instrumented; device code unchanged) into
# $Q3ASR_SAN_DIR/{address,undefined}/ (default: a scratch directory under /tmp) and runs the host-only tests and tools/fuzz_host.py against each.  No GPU needed.
# Usage: tools/sanitize.sh [iterations]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
PKG=$ROOT/qwen3-asr-swift_b200
IT=${1:-800}
OUT=${Q3ASR_SAN_DIR:-/tmp/q3asr_san}
TESTS="tests/test_tokenizer.py tests/test_audio_io.py tests/test_aligner.py tests/test_checkpoint_index.py tests/test_prompt.py tests/test_text_preprocessing.py tests/test_abi.py tests/test_sampler.py"
mkdir -p $OUT/site
cat > $OUT/site/sitecustomize.py <<PY
import os, sys
if os.environ.get("Q3LIB"):
    sys.path.insert(0, "$PKG")
    import q3asr
    q3asr.LIB_PATH = os.environ["Q3LIB"]
PY
for san in address undefined thread; do
  d=$OUT/$san; mkdir -p $d
  for f in $PKG/csrc/*.cu; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -Xcompiler -fPIC,-fsanitize=$san,-fno-omit-frame-pointer \
         -I$ROOT/include -I$PKG/csrc --expt-relaxed-constexpr -c $f -o $d/$(basename $f .cu).o &
  done; wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $d/libq3asr.so $d/*.o -lcudart_static -lpthread -ldl -lrt -Xcompiler -fsanitize=$san
  case $san in address) rt=libasan.so;; undefined) rt=libubsan.so;; thread) rt=libtsan.so;; esac
  rt=$(gcc -print-file-name=$rt)
  # libstdc++ is preloaded too: the sanitizer runtime intercepts __cxa_throw and must find the real one at start-up
  export LD_PRELOAD="$rt /usr/lib/x86_64-linux-gnu/libstdc++.so.6" ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1
  export Q3LIB=$d/libq3asr.so PYTHONPATH=$OUT/site
  if [ $san = thread ]; then  # ThreadSanitizer: the eight-thread re-entrancy test only (any report is printed as "WARNING: ThreadSanitizer")
    (cd $ROOT && TSAN_OPTIONS="report_signal_unsafe=0" python -m pytest tests/test_abi.py -q -k reentrant -p no:cacheprovider 2>&1 | grep -E "ThreadSanitizer|passed|failed")
  else
    (cd $ROOT && python -m pytest $TESTS -x -q -m "not gpu" -p no:cacheprovider | tail -2)
    python $ROOT/tools/fuzz_host.py 1 $d/scratch $IT
  fi
  unset LD_PRELOAD
done
