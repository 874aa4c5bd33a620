#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
#
# Rebuilds the reference's own kernels, wrappers and launch-configuration resolver (eyalroz/libgiddy @
# 74d3543, from the sources where they lie under /root/reference) for sm_100a, behind the same
# registry C ABI as the product (libgiddy_b200/csrc/host/registry_c_abi.cpp compiled against the
# reference's headers).  Output: oracle/_ref/libgiddy_ref.so (git-ignored; travels to the GPU box).
#
# The reference does not build as-is with CUDA 12.9 / gcc 13 and has no Boost here; the sources are
# copied to a scratch directory OUTSIDE the repository and patched there, mechanically (SURVEY.md
# Appendix D).  No reference source is copied into the repository.  Its own CMake build is not used.
set -euo pipefail

REPO="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
REFERENCE="${GIDDY_REFERENCE:-/root/reference}"
SCRATCH="${GIDDY_REF_SCRATCH:-/tmp/giddy_ref_build}"
OUT="$REPO/oracle/_ref"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
CXX=/usr/bin/g++
JOBS="${JOBS:-8}"

if [ ! -d "$REFERENCE/src" ]; then echo "no reference at $REFERENCE: nothing to build"; exit 0; fi
if [ -f "$OUT/libgiddy_ref.so" ] && [ "$OUT/libgiddy_ref.so" -nt "$REPO/libgiddy_b200/csrc/host/registry_c_abi.cpp" ] \
   && [ "$OUT/libgiddy_ref.so" -nt "${BASH_SOURCE[0]}" ]; then
	echo "oracle/_ref/libgiddy_ref.so is up to date"; exit 0
fi

rm -rf "$SCRATCH"; mkdir -p "$SCRATCH/shim/boost/optional" "$SCRATCH/obj" "$OUT"
cp -rL "$REFERENCE/src" "$SCRATCH/src"    # -L: src/cuda/api and api_wrappers.h are symlinks into external/
chmod -R u+w "$SCRATCH/src"
S="$SCRATCH/src"

# 1. inside the reference's namespace cuda, unqualified std resolves to libcu++'s cuda::std
find "$S" -type f | xargs sed -i -E 's/([^:A-Za-z_]|^)std::/\1::std::/g'
# 2. cudaPointerAttributes::memoryType was renamed
sed -i 's/attributes\.memoryType/attributes.type/g' "$S/cuda/api/memory.hpp"
# 3. legacy warp votes
sed -i -E 's/\b__ballot\(([^)]*)\)/__ballot_sync(0xffffffffu, \1)/; s/\b__all\(([^)]*)\)/__all_sync(0xffffffffu, \1)/; s/\b__any\(([^)]*)\)/__any_sync(0xffffffffu, \1)/' "$S/cuda/on_device/builtins.cuh"
# 4. legacy shuffles used by the generic shuffle templates
python3 - "$S/cuda/on_device/generic_shuffle.cuh" <<'PY'
import sys
p = sys.argv[1]; s = open(p).read()
marker = '#include <sm_30_intrinsics.h>'
shims = marker + '''
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 700
__device__ __forceinline__ int __shfl     (int v, int lane, int w)   { return __shfl_sync     (0xffffffffu, v, lane, w); }
__device__ __forceinline__ int __shfl_up  (int v, unsigned d, int w) { return __shfl_up_sync  (0xffffffffu, v, d, w); }
__device__ __forceinline__ int __shfl_down(int v, unsigned d, int w) { return __shfl_down_sync(0xffffffffu, v, d, w); }
__device__ __forceinline__ int __shfl_xor (int v, int m, int w)      { return __shfl_xor_sync (0xffffffffu, v, m, w); }
#endif
'''
assert marker in s
open(p, 'w').write(s.replace(marker, shims, 1))
PY
# 5. 64-bit special registers / bfind with 32-bit asm constraints
python3 - "$S/cuda/on_device/ptx.cuh" <<'PY'
import re, sys
p = sys.argv[1]; s = open(p).read()
repl = {
 r'DEFINE_SPECIAL_REGISTER_GETTER\(\s*gridid\s*,\s*uint64_t\s*,\s*64\s*\)': '__forceinline__ __device__ uint64_t gridid()  { uint64_t r; asm("mov.u64 %0, %gridid;"  : "=l"(r)); return r; }',
 r'DEFINE_SPECIAL_REGISTER_GETTER\(\s*clock64\s*,\s*uint64_t\s*,\s*64\s*\)': '__forceinline__ __device__ uint64_t clock64() { uint64_t r; asm("mov.u64 %0, %clock64;" : "=l"(r)); return r; }',
 r'DEFINE_BFIND\(\s*int64_t\s*,\s*64\s*\)': '__forceinline__ __device__ uint32_t bfind(int64_t v)  { uint32_t r; asm("bfind.s64 %0, %1;" : "=r"(r) : "l"(v)); return r; }',
 r'DEFINE_BFIND\(\s*uint64_t\s*,\s*64\s*\)': '__forceinline__ __device__ uint32_t bfind(uint64_t v) { uint32_t r; asm("bfind.u64 %0, %1;" : "=r"(r) : "l"(v)); return r; }',
}
for pattern, text in repl.items():
    s, count = re.subn(pattern + r'\s*;?', text, s)
    assert count >= 1, pattern
open(p, 'w').write(s)
PY
# 6. missing include
sed -i '0,/#include/s//#include <limits>\n#include/' "$S/util/integer.h"
# 7. capability tables end at compute capability 6.2: add rows for 10.0 (B200)
sed -i -E 's/^(\s*)\{ 60, 48 \* 1024\s*\},/&\n\1{ 100, 48 * 1024 },/; s/^(\s*)\{ 60, 64 \},/&\n\1{ 100, 64 },/; s/^(\s*)\{ 50, 4 \},/&\n\1{ 100, 4 },/' "$S/cuda/api/device_properties.cpp"
grep -c "{ 100," "$S/cuda/api/device_properties.cpp" | grep -qx 3 || { echo "device_properties.cpp: capability-table patch did not apply"; exit 1; }

# Boost shims (any / optional / none / integer), C++17 equivalents
cat > "$SCRATCH/shim/boost/any.hpp" <<'H'
#pragma once
#include <any>
namespace boost { using any = ::std::any; using ::std::any_cast; using bad_any_cast = ::std::bad_any_cast; }
H
cat > "$SCRATCH/shim/boost/optional.hpp" <<'H'
#pragma once
#include <optional>
#include <boost/none_t.hpp>
namespace boost { template <class T> using optional = ::std::optional<T>; }
H
cp "$SCRATCH/shim/boost/optional.hpp" "$SCRATCH/shim/boost/optional/optional_io.hpp"
cat > "$SCRATCH/shim/boost/none_t.hpp" <<'H'
#pragma once
#include <optional>
namespace boost { using none_t = ::std::nullopt_t; inline constexpr ::std::nullopt_t none = ::std::nullopt; }
H
cat > "$SCRATCH/shim/boost/integer.hpp" <<'H'
#pragma once
#include <cstdint>
namespace boost {
template <int Bits> struct uint_t {
	using least = typename ::std::conditional<(Bits <= 8), ::std::uint8_t, typename ::std::conditional<(Bits <= 16), ::std::uint16_t,
		typename ::std::conditional<(Bits <= 32), ::std::uint32_t, ::std::uint64_t>::type>::type>::type;
	using fast = least; using exact = least;
};
template <int Bits> struct int_t {
	using least = typename ::std::conditional<(Bits <= 8), ::std::int8_t, typename ::std::conditional<(Bits <= 16), ::std::int16_t,
		typename ::std::conditional<(Bits <= 32), ::std::int32_t, ::std::int64_t>::type>::type>::type;
	using fast = least; using exact = least;
};
}
H
sed -i '1i #include <type_traits>' "$SCRATCH/shim/boost/integer.hpp"

INC="-I$S -I$SCRATCH/shim -I$REPO/include"
NVFLAGS="-w -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $INC"

compile_cu() { "$NVCC" $NVFLAGS -c "$1" -o "$2" 2> "$2.log" || { tail -30 "$2.log"; echo "FAILED: $1"; return 1; }; }
export -f compile_cu; export NVCC NVFLAGS

pids=()
run_limited() { while [ "$(jobs -rp | wc -l)" -ge "$JOBS" ]; do sleep 0.2; done; "$@" & pids+=($!); }
for f in "$S"/kernel_wrappers/decompression/*.cu "$S"/kernel_wrappers/data_layout/*.cu; do
	run_limited compile_cu "$f" "$SCRATCH/obj/ref_$(basename "$f" .cu).o"
done
# the registry C ABI, compiled against the reference's headers, and the error/query half of the C ABI
run_limited "$NVCC" $NVFLAGS -x cu -DGIDDY_REGISTRY_FOR_REFERENCE -c "$REPO/libgiddy_b200/csrc/host/registry_c_abi.cpp" -o "$SCRATCH/obj/registry_c_abi.o"
run_limited "$NVCC" $NVFLAGS -c "$REPO/libgiddy_b200/csrc/abi/abi_common.cu" -o "$SCRATCH/obj/abi_common.o"
run_limited "$CXX" -w -std=c++17 -O2 -fPIC $INC -I/usr/local/cuda/include -c "$S/kernels/resolve_launch_configuration.cpp" -o "$SCRATCH/obj/resolve_launch_configuration.o"
run_limited "$CXX" -w -std=c++17 -O2 -fPIC $INC -I/usr/local/cuda/include -c "$S/cuda/api/device_properties.cpp" -o "$SCRATCH/obj/device_properties.o"
status=0
for pid in "${pids[@]}"; do wait "$pid" || status=1; done
[ "$status" -eq 0 ] || { echo "reference build failed"; exit 1; }

"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libgiddy_ref.so" "$SCRATCH"/obj/*.o -cudart static
echo "built $OUT/libgiddy_ref.so from $REFERENCE (scratch: $SCRATCH)"
rm -rf "$SCRATCH/src"   # keep no copy of the reference's sources around
