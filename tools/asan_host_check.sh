#!/bin/bash
# AddressSanitizer (+ UndefinedBehaviorSanitizer with SAN="address undefined") build of the HOST side of the library — engine,
# schedulers, fusion, the launchers' argument checks and parameter-fill functions, the C ABI, the pybind modules — in a scratch
# copy (nothing in the tree is touched), then the CPU tests of that code and the invalid-input fuzzers under it.  No GPU:
# device code is compiled as usual and never runs.    tools/asan_host_check.sh [scratch dir]   (about 10 min on 8 cores)
#
# Python is not linked against libstdc++, so libasan's __cxa_throw interceptor finds no real function unless libstdc++ is
# preloaded next to it (otherwise: "CHECK failed: asan_interceptors.cpp ... real___cxa_throw != 0" at the first C++ throw).
set -eu
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=${1:-/tmp/hiq_asan}
SAN=${SAN:-address}
rm -rf "$W" && mkdir -p "$W"
cp -r "$ROOT"/hiqsimulator_b200 "$ROOT"/include "$ROOT"/tests "$ROOT"/oracle "$ROOT"/tools "$ROOT"/hiq "$ROOT"/bench.py "$ROOT"/bench_extra.py \
      "$ROOT"/__graft_entry__.py "$W"/
cd "$W"/hiqsimulator_b200/csrc
rm -rf build ../*.so
GXX=""; NVX=""
for s in $SAN; do GXX="$GXX -fsanitize=$s"; NVX="$NVX,-fsanitize=$s"; done   # nvcc's -Xcompiler list is comma-separated
sed -i "s/-Xcompiler -fPIC,-Wall,-Wno-unknown-pragmas/-Xcompiler -fPIC,-Wall,-Wno-unknown-pragmas$NVX,-fno-omit-frame-pointer,-g/;
        s/^CXXFLAGS := -O2/CXXFLAGS := -O1 -g$GXX -fno-omit-frame-pointer/;
        s/-ldl -Xlinker/-ldl -Xcompiler ${NVX#,} -Xlinker/" Makefile
make -j"$(nproc)" > "$W"/build.log 2>&1 || { tail -n 20 "$W"/build.log; exit 1; }
cd "$W"
GCCDIR=$(dirname "$(gcc -print-file-name=libasan.so)")
PRE="$GCCDIR/libasan.so $(gcc -print-file-name=libstdc++.so.6)"
case "$SAN" in *undefined*) PRE="$PRE $GCCDIR/libubsan.so";; esac
export ASAN_OPTIONS=detect_leaks=0:abort_on_error=1:halt_on_error=1
export UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1
run() { echo "== $*"; LD_PRELOAD="$PRE" "$@" 2>&1 | grep -v "^  File\|^    #[0-9]* 0x[0-9a-f]* .*python3" | tail -n 6; }
run python -m pytest tests/test_host_logic.py tests/test_abi.py tests/test_scheduler.py tests/test_tile_program_cpu.py \
    tests/test_diag_program_cpu.py tests/test_dense_launcher_cpu.py tests/test_edge_cases_cpu.py -q -p no:cacheprovider
run python tools/fuzz_launcher_arguments.py 0 4000
run python tools/fuzz_sched_invalid.py 0 2000
run python tools/fuzz_invalid_arguments.py 0 500
run python tools/fuzz_invalid_arguments.py 0 500 --ranks
run python tools/fuzz_launch_trace.py 0 40
echo "sanitizer reports: $(grep -l 'ERROR: AddressSanitizer\|runtime error' "$W"/*.log 2>/dev/null | wc -l) (none expected; output above)"
