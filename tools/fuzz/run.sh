#!/bin/bash
# Memory-safety check of the decoder's SOURCE (host parser + the per-thread device functions) on mutated JPEG files:
# the CPU harness (tests/emu/emu_driver.cpp: emu_decode) built with AddressSanitizer.  No GPU involved.
#   tools/fuzz/run.sh [workdir]        (4000 mutants of five seed files; prints "decoded N rejected M" per 500 files)
set -e
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
W="${1:-/tmp/jpeg_gpu_fuzz}"
mkdir -p "$W" && cd "$W"
sed "s#/tmp/fz/in#$W/in#g; s#/root/repo#$ROOT#g" "$ROOT/tools/fuzz/gen_mutants.py" > gen.py
rm -rf in && PYTHONPATH="$ROOT" python gen.py
g++ -std=c++17 -O1 -g -fsanitize=address -fno-omit-frame-pointer -pthread -I"$ROOT/tests/emu" -I"$ROOT/imagecodecs_b200/csrc" -o fuzz_dec \
    "$ROOT/tools/fuzz/fuzz_decoder_main.cpp" "$ROOT/tests/emu/emu_driver.cpp" "$ROOT/imagecodecs_b200/csrc/jpeg_host.cpp" "$ROOT/imagecodecs_b200/csrc/jpeg_decode_host.cpp"
ls in/*.jpg | xargs -n 500 ./fuzz_dec
# the encode kernel source under AddressSanitizer (18 shapes: plain / two-iteration / restart kernels, slow path, swizzles)
g++ -std=c++17 -O1 -g -fsanitize=address -fno-omit-frame-pointer -ffp-contract=off -pthread -I"$ROOT/tests/emu" -I"$ROOT/imagecodecs_b200/csrc" -o asan_enc \
    "$ROOT/tools/fuzz/asan_encoder_main.cpp" "$ROOT/tests/emu/emu_driver.cpp" "$ROOT/imagecodecs_b200/csrc/jpeg_host.cpp" "$ROOT/imagecodecs_b200/csrc/jpeg_decode_host.cpp"
./asan_enc | tail -2
# the subsequence rounds with eight racing host threads per round under ThreadSanitizer (records are updated in place)
g++ -std=c++17 -O1 -g -fsanitize=thread -pthread -I"$ROOT/tests/emu" -I"$ROOT/imagecodecs_b200/csrc" -o tsan_dec \
    "$ROOT/tools/fuzz/tsan_rounds_main.cpp" "$ROOT/tests/emu/emu_driver.cpp" "$ROOT/imagecodecs_b200/csrc/jpeg_host.cpp" "$ROOT/imagecodecs_b200/csrc/jpeg_decode_host.cpp"
PYTHONPATH="$ROOT" python -c "
import oracle
open('t1.jpg','wb').write(oracle.oracle_encode(oracle.synth_image(640,360,3),1,75,1))
open('t2.jpg','wb').write(oracle.oracle_encode(oracle.synth_image(320,200,3),0,3,0))"
./tsan_dec t1.jpg t2.jpg && echo "no race report"
