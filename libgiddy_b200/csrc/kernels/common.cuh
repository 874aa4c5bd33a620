// Device-side building blocks shared by all giddy_b200 kernels (sm_100a).
//
// Everything here is integer-only.  The recurring ideas:
//   * every global access a thread makes is as wide as alignment allows (up to
//     128 bits) and warp-contiguous, so each warp instruction touches whole
//     32-byte sectors;
//   * decoded output is never re-read by these kernels, so it is written with
//     streaming (evict-first) stores; compressed input is read through the
//     non-coherent path without L1 allocation unless it is going to be re-read;
//   * work is distributed at warp granularity with grid-stride loops, so every
//     kernel is correct for any grid size and any block size that is a multiple
//     of 32 -- the launch configuration resolver only affects speed.
#pragma once
#ifndef GIDDY_B200_KERNELS_COMMON_CUH_
#define GIDDY_B200_KERNELS_COMMON_CUH_

#include <cstdint>
#include <cstddef>
#include <type_traits>
#include <cuda_runtime.h>
#include "giddy/kernel_traits.h"

namespace giddy {

template <unsigned N> struct uint_of;
template <> struct uint_of<1> { using type = uint8_t;  };
template <> struct uint_of<2> { using type = uint16_t; };
template <> struct uint_of<4> { using type = uint32_t; };
template <> struct uint_of<8> { using type = uint64_t; };
template <unsigned N> using uint_t = typename uint_of<N>::type;

template <unsigned N> struct int_of;
template <> struct int_of<1> { using type = int8_t;  };
template <> struct int_of<2> { using type = int16_t; };
template <> struct int_of<4> { using type = int32_t; };
template <> struct int_of<8> { using type = int64_t; };
template <unsigned N> using int_t = typename int_of<N>::type;

// size_type_by_index_size<N> of the reference (src/kernels/common.cuh:49-53):
// 32 bits for index sizes up to 4, 64 bits for 8.
template <unsigned IndexSize>
using size_type = typename std::conditional<(IndexSize > 4), uint64_t, uint32_t>::type;

// POD array passed by value / stored per segment: the layout of the reference's
// cuda::array<T, N> (src/cuda/stl_array.cuh:15-18,63).
template <typename T, unsigned N> struct pod_array { T v[N]; };

enum : int { model_constant = 0, model_linear = 1 };

namespace device {

constexpr unsigned warp_size = 32;
constexpr unsigned full_mask = 0xffffffffu;

__device__ __forceinline__ unsigned lane_index() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned warp_index_in_block() { return threadIdx.x >> 5; }
__device__ __forceinline__ unsigned warps_per_block() { return blockDim.x >> 5; }
__device__ __forceinline__ uint64_t global_warp_index() { return (uint64_t) blockIdx.x * warps_per_block() + warp_index_in_block(); }
__device__ __forceinline__ uint64_t num_warps_in_grid() { return (uint64_t) gridDim.x * warps_per_block(); }

// bytes must be a power of 2
__device__ __forceinline__ bool is_aligned_to(const void* p, unsigned bytes)
{
	return (reinterpret_cast<uintptr_t>(p) & (uintptr_t)(bytes - 1)) == 0;
}

// ---- 128-bit streaming accessors -------------------------------------------------------------

__device__ __forceinline__ uint4 load_streaming_16(const void* p)
{
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
		: "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}

__device__ __forceinline__ uint2 load_streaming_8(const void* p)
{
	uint2 r;
	asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
	return r;
}

__device__ __forceinline__ uint32_t load_streaming_4(const void* p)
{
	uint32_t r;
	asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
	return r;
}

__device__ __forceinline__ uint16_t load_streaming_2(const void* p)
{
	uint16_t r;
	asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(p));
	return r;
}

__device__ __forceinline__ uint8_t load_streaming_1(const void* p)
{
	uint32_t r;
	asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(r) : "l"(p));
	return (uint8_t) r;
}

__device__ __forceinline__ void store_streaming_16(void* p, const uint4& v)
{
	asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// 256-bit streaming store (sm_100: STG.E.EF.ENL2.256); p must be 32-byte aligned
__device__ __forceinline__ void store_streaming_32(void* p, const uint32_t (&w)[8])
{
	asm volatile("st.global.cs.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
		:: "l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}

// A register-resident group of Bytes bytes, loadable in one (or a few) maximal-width
// streaming accesses and addressable as elements of any power-of-2 size.
template <unsigned Bytes>
struct alignas(Bytes >= 16 ? 16 : Bytes) byte_pack {
	static_assert(Bytes == 1 || Bytes == 2 || Bytes == 4 || Bytes == 8 || Bytes % 16 == 0, "unsupported pack size");
	union {
		uint8_t  u8[Bytes];
		uint16_t u16[(Bytes + 1) / 2];
		uint32_t u32[(Bytes + 3) / 4];
		uint64_t u64[(Bytes + 7) / 8];
	};

	template <unsigned ElementSize>
	__device__ __forceinline__ uint_t<ElementSize> get(unsigned i) const
	{
		if (ElementSize == 1) { return (uint_t<ElementSize>) u8[i]; }
		if (ElementSize == 2) { return (uint_t<ElementSize>) u16[i]; }
		if (ElementSize == 4) { return (uint_t<ElementSize>) u32[i]; }
		return (uint_t<ElementSize>) u64[i];
	}

	template <unsigned ElementSize>
	__device__ __forceinline__ void set(unsigned i, uint_t<ElementSize> x)
	{
		if (ElementSize == 1) { u8[i] = (uint8_t) x; }
		else if (ElementSize == 2) { u16[i] = (uint16_t) x; }
		else if (ElementSize == 4) { u32[i] = (uint32_t) x; }
		else { u64[i] = (uint64_t) x; }
	}
};

// Load Bytes bytes from a Bytes-aligned (16-aligned when Bytes >= 16) global address.
template <unsigned Bytes>
__device__ __forceinline__ byte_pack<Bytes> load_pack(const void* p)
{
	byte_pack<Bytes> r;
	if (Bytes == 1) { r.u8[0] = load_streaming_1(p); }
	else if (Bytes == 2) { r.u16[0] = load_streaming_2(p); }
	else if (Bytes == 4) { r.u32[0] = load_streaming_4(p); }
	else if (Bytes == 8) { uint2 t = load_streaming_8(p); r.u32[0] = t.x; r.u32[1] = t.y; }
	else {
		#pragma unroll
		for (unsigned k = 0; k < Bytes / 16; k++) {
			uint4 t = load_streaming_16(static_cast<const uint8_t*>(p) + 16 * k);
			r.u32[4 * k] = t.x; r.u32[4 * k + 1] = t.y; r.u32[4 * k + 2] = t.z; r.u32[4 * k + 3] = t.w;
		}
	}
	return r;
}

__device__ __forceinline__ void store_pack_16(void* p, const byte_pack<16>& v)
{
	store_streaming_16(p, make_uint4(v.u32[0], v.u32[1], v.u32[2], v.u32[3]));
}

// ---- TMA bulk copies (global -> shared) completing on an mbarrier -------------------------------
// One thread arms the barrier with the byte count and issues the copy (SASS: UBLKCP); every thread
// that will read the staged bytes waits for the barrier's phase to flip.  Source, destination and size
// must be multiples of 16 bytes.

__device__ __forceinline__ uint32_t shared_address(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbarrier_init(uint64_t* barrier, unsigned arrivals)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(shared_address(barrier)), "r"(arrivals) : "memory");
	// make the initialised barrier visible to the async (TMA) proxy
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbarrier_expect_bytes(uint64_t* barrier, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(shared_address(barrier)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbarrier_arrive(uint64_t* barrier)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(shared_address(barrier)) : "memory");
}

__device__ __forceinline__ void mbarrier_wait(uint64_t* barrier, uint32_t phase_parity)
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_LOOP:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra WAIT_DONE;\n"
		"bra WAIT_LOOP;\n"
		"WAIT_DONE:\n"
		"}\n" :: "r"(shared_address(barrier)), "r"(phase_parity) : "memory");
}

// Order this thread's earlier generic-proxy accesses to shared memory before later async-proxy (TMA) ones.
__device__ __forceinline__ void fence_proxy_async_shared()
{
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void tma_bulk_load(void* shared_destination, const void* global_source, uint32_t bytes, uint64_t* barrier)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		:: "r"(shared_address(shared_destination)), "l"(global_source), "r"(bytes), "r"(shared_address(barrier)) : "memory");
}

// ---- programmatic dependent launch ---------------------------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while the kernel before it on the
// stream is still running -- as soon as all of that kernel's CTAs have called launch_dependent_grids() (or exited) -- and
// must call wait_for_prerequisite_grids() before it touches anything that kernel writes: the call returns once the earlier
// kernel has completed and its writes are visible.  Both are no-ops in a kernel launched the ordinary way.
__device__ __forceinline__ void launch_dependent_grids() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void wait_for_prerequisite_grids() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ uint32_t dynamic_shared_memory_size()
{
	uint32_t r;
	asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(r));
	return r;
}

// ---- warp-level scans ------------------------------------------------------------------------

template <typename T>
__device__ __forceinline__ T shuffle_up(T x, unsigned delta)
{
	if (sizeof(T) <= 4) { return (T) __shfl_up_sync(full_mask, (uint32_t) x, delta); }
	else { return (T) __shfl_up_sync(full_mask, (unsigned long long) x, delta); }
}

template <typename T>
__device__ __forceinline__ T shuffle_from(T x, unsigned lane)
{
	if (sizeof(T) <= 4) { return (T) __shfl_sync(full_mask, (uint32_t) x, lane); }
	else { return (T) __shfl_sync(full_mask, (unsigned long long) x, lane); }
}

// Inclusive prefix sum across the warp's lanes (modular arithmetic in T).
template <typename T>
__device__ __forceinline__ T warp_inclusive_sum(T x)
{
	#pragma unroll
	for (unsigned d = 1; d < warp_size; d <<= 1) {
		T y = shuffle_up(x, d);
		if (lane_index() >= d) { x = (T)(x + y); }
	}
	return x;
}

// ---- integer helpers --------------------------------------------------------------------------

template <typename T>
__host__ __device__ constexpr T div_rounding_up(T a, T b) { return (a + b - 1) / b; }

template <typename T>
__device__ __forceinline__ T minimum(T a, T b) { return a < b ? a : b; }

// Per-warp slice of the block's dynamic shared memory.
extern __shared__ __align__(16) unsigned char dynamic_shared_memory[];

__device__ __forceinline__ unsigned char* warp_shared_memory(unsigned bytes_per_warp)
{
	return dynamic_shared_memory + (size_t) warp_index_in_block() * bytes_per_warp;
}

} // namespace device
} // namespace giddy

#endif // GIDDY_B200_KERNELS_COMMON_CUH_
