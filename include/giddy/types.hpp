// giddy/types.hpp -- the plain types which appear in the signatures of the kernel-wrapper API.
//
// These mirror, name for name, the handful of types the reference takes from its
// cuda-api-wrappers submodule (external/cuda-api-wrappers/src/cuda/api/types.h:40-190,
// device_function.hpp, device_properties.hpp) and from src/cuda/optional_and_any.hpp, so that code
// written against the reference's wrapper API compiles against this library.  They live in
// namespace cuda, as in the reference; every standard-library name in these headers is written
// ::std:: so that they coexist with libcu++'s cuda::std.
//
// Host-only: compiles with g++ (needs only cuda_runtime_api.h for the CUDA struct definitions).
#pragma once
#ifndef GIDDY_TYPES_HPP_
#define GIDDY_TYPES_HPP_

#include "giddy_b200.h"

#include <cuda_runtime_api.h>

#include <any>
#include <optional>
#include <cstdint>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <type_traits>

namespace util {

enum class endianness_t : bool { big_endian = false, little_endian = true, big = big_endian, little = little_endian };

template <unsigned N> struct uint_by_size;
template <> struct uint_by_size<1> { using type = ::std::uint8_t;  };
template <> struct uint_by_size<2> { using type = ::std::uint16_t; };
template <> struct uint_by_size<4> { using type = ::std::uint32_t; };
template <> struct uint_by_size<8> { using type = ::std::uint64_t; };

/// An unsigned integer of exactly N bytes (reference: src/util/integer.h:203-210)
template <unsigned N> using uint_t = typename uint_by_size<N>::type;

template <typename T> constexpr T div_rounding_up(T a, T b) { return (a + b - 1) / b; }
template <typename T> constexpr T round_down(T x, T y) { return x - x % y; }
template <typename T> constexpr T round_up(T x, T y) { return div_rounding_up(x, y) * y; }

} // namespace util

namespace cuda {

// ---- optional / any (reference: src/cuda/optional_and_any.hpp:14-68; Boost there, std here) -----

template <typename T> using optional = ::std::optional<T>;
using any = ::std::any;
using bad_any_cast = ::std::bad_any_cast;
inline constexpr ::std::nullopt_t nullopt = ::std::nullopt;

template <typename T>
inline T any_cast(const any& operand) { return ::std::any_cast<T>(operand); }

template <typename T>
inline optional<T> any_cast(const optional<any>& operand)
{
	return operand ? optional<T>(::std::any_cast<T>(operand.value())) : optional<T>(nullopt);
}

template <typename C, typename E>
inline optional<typename C::mapped_type> maybe_at(const C& container, const E& element)
{
	auto it = container.find(element);
	return (it == container.end()) ? optional<typename C::mapped_type>(nullopt) : optional<typename C::mapped_type>(it->second);
}

// ---- scalars ------------------------------------------------------------------------------------

using native_word_t          = unsigned;
using grid_dimension_t       = unsigned;
using grid_block_dimension_t = unsigned;
using serialization_factor_t = unsigned short;

/// Widened from the reference's unsigned short (types.h:93): a B200 block may use 227 KiB.
using shared_memory_size_t   = unsigned;

enum : native_word_t { warp_size = 32, log_warp_size = 5 };

enum class endianness_t : bool { big, big_endian = big, little, little_endian = little };

using util::uint_t;

/// reference: src/cuda/miscellany.h:29-30
template <typename Size> using promoted_size = typename ::std::common_type<Size, native_word_t>::type;
/// reference: src/kernels/common.cuh:49-53
template <unsigned N> using size_type_by_index_size = promoted_size<uint_t<N>>;
template <typename T> using size_type_by_index_type = size_type_by_index_size<sizeof(T)>;

/// A POD fixed-size array usable as a kernel parameter (reference: src/cuda/stl_array.cuh:15-63)
template <typename T, ::std::size_t N>
struct array {
	using value_type = T;
	T elements_[N ? N : 1];
	constexpr T& operator[](::std::size_t i) { return elements_[i]; }
	constexpr const T& operator[](::std::size_t i) const { return elements_[i]; }
	constexpr ::std::size_t size() const { return N; }
	constexpr T* data() { return elements_; }
	constexpr const T* data() const { return elements_; }
	constexpr T* begin() { return elements_; }
	constexpr T* end() { return elements_ + N; }
	constexpr const T* begin() const { return elements_; }
	constexpr const T* end() const { return elements_ + N; }
};

// ---- launch configuration -----------------------------------------------------------------------

struct dimensions_t {
	unsigned x, y, z;
	constexpr dimensions_t(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) { }
	constexpr ::std::size_t volume() const { return (::std::size_t) x * y * z; }
	constexpr bool empty() const { return volume() == 0; }
};
constexpr inline bool operator==(const dimensions_t& l, const dimensions_t& r) { return l.x == r.x && l.y == r.y && l.z == r.z; }

using grid_dimensions_t       = dimensions_t;
using grid_block_dimensions_t = dimensions_t;

struct launch_configuration_t {
	grid_dimensions_t       grid_dimensions;
	grid_block_dimensions_t block_dimensions;
	shared_memory_size_t    dynamic_shared_memory_size; // in bytes
};

inline launch_configuration_t make_launch_config(
	grid_dimensions_t grid_dimensions, grid_block_dimensions_t block_dimensions, shared_memory_size_t dynamic_shared_memory_size = 0)
{
	return { grid_dimensions, block_dimensions, dynamic_shared_memory_size };
}

inline bool operator==(const launch_configuration_t& l, const launch_configuration_t& r)
{
	return l.grid_dimensions == r.grid_dimensions && l.block_dimensions == r.block_dimensions &&
		l.dynamic_shared_memory_size == r.dynamic_shared_memory_size;
}

inline giddy_launch_config_t to_c(const launch_configuration_t& c)
{
	return { { c.grid_dimensions.x, c.grid_dimensions.y, c.grid_dimensions.z },
	         { c.block_dimensions.x, c.block_dimensions.y, c.block_dimensions.z }, c.dynamic_shared_memory_size };
}

// ---- errors ---------------------------------------------------------------------------------------

/// Thrown when a CUDA runtime call made on the caller's behalf fails (reference: external/.../error.hpp)
class runtime_error : public ::std::runtime_error {
public:
	using ::std::runtime_error::runtime_error;
};

namespace detail {

/// Turn a failed C-ABI call into the exception the reference would have thrown.
[[noreturn]] inline void throw_for(giddy_status_t status)
{
	::std::string message = giddy_last_error();
	switch (status) {
	case GIDDY_INVALID_ARGUMENT: throw ::std::invalid_argument(message);
	case GIDDY_OUT_OF_RANGE:     throw ::std::out_of_range(message);
	case GIDDY_LOGIC_ERROR:      throw ::std::logic_error(message);
	case GIDDY_BAD_ANY_CAST:     throw bad_any_cast();
	case GIDDY_UNSUPPORTED:      throw ::std::invalid_argument(message);
	default:                     throw runtime_error(message);
	}
}

inline void throw_if_error(giddy_status_t status) { if (status != GIDDY_SUCCESS) { throw_for(status); } }

} // namespace detail

// ---- streams, devices, device functions -----------------------------------------------------------

namespace stream {
using id_t = cudaStream_t;
const id_t default_stream_id = nullptr;
} // namespace stream

namespace device {

using id_t = int;

struct compute_capability_t {
	unsigned major, minor;
	constexpr unsigned as_combined_number() const { return major * 10 + minor; }
	/// warp schedulers per SM; 4 on every architecture since Maxwell, including Blackwell
	constexpr unsigned max_warp_schedulings_per_processor_cycle() const { return (major == 3) ? 4 : (major < 3 ? 2 : 4); }
};

struct properties_t : public cudaDeviceProp {
	properties_t() = default;
	properties_t(const cudaDeviceProp& p) : cudaDeviceProp(p) { }
	compute_capability_t compute_capability() const { return { (unsigned) major, (unsigned) minor }; }
	grid_block_dimension_t max_threads_per_block() const { return (grid_block_dimension_t) maxThreadsPerBlock; }
	grid_block_dimension_t max_warps_per_block() const { return max_threads_per_block() / warp_size; }
};

/// Properties of a device, via the library (so that no CUDA runtime needs to be linked by the caller)
inline properties_t get_properties(id_t device_id = -1)
{
	giddy_device_properties_t p;
	detail::throw_if_error(giddy_device_properties(device_id, &p));
	properties_t r;
	r = properties_t(cudaDeviceProp{});
	r.major = p.compute_capability_major;
	r.minor = p.compute_capability_minor;
	r.multiProcessorCount = p.multiprocessor_count;
	r.maxThreadsPerBlock = p.max_threads_per_block;
	r.maxThreadsDim[0] = p.max_threads_dim_x;
	r.maxThreadsDim[1] = r.maxThreadsDim[2] = 1;
	r.maxThreadsPerMultiProcessor = p.max_threads_per_multiprocessor;
	r.maxBlocksPerMultiProcessor = p.max_blocks_per_multiprocessor;
	r.warpSize = p.warp_size;
	r.sharedMemPerBlock = p.shared_memory_per_block;
	r.sharedMemPerBlockOptin = p.shared_memory_per_block_optin;
	r.sharedMemPerMultiprocessor = p.shared_memory_per_multiprocessor;
	r.l2CacheSize = (int) p.l2_cache_size;
	r.totalGlobalMem = p.total_global_memory;
	return r;
}

namespace current {
struct handle_t { properties_t properties() const { return get_properties(-1); } };
inline handle_t get() { return { }; }
} // namespace current

inline giddy_device_properties_t to_c(const properties_t& p)
{
	giddy_device_properties_t r;
	r.compute_capability_major = p.major;
	r.compute_capability_minor = p.minor;
	r.multiprocessor_count = p.multiProcessorCount;
	r.max_threads_per_block = p.maxThreadsPerBlock;
	r.max_threads_dim_x = p.maxThreadsDim[0];
	r.max_threads_per_multiprocessor = p.maxThreadsPerMultiProcessor;
	r.max_blocks_per_multiprocessor = p.maxBlocksPerMultiProcessor;
	r.warp_size = p.warpSize;
	r.shared_memory_per_block = p.sharedMemPerBlock;
	r.shared_memory_per_block_optin = p.sharedMemPerBlockOptin;
	r.shared_memory_per_multiprocessor = p.sharedMemPerMultiprocessor;
	r.l2_cache_size = (::std::size_t) p.l2CacheSize;
	r.total_global_memory = p.totalGlobalMem;
	return r;
}

} // namespace device

namespace device_function {

struct attributes_t : public cudaFuncAttributes {
	attributes_t() : cudaFuncAttributes{} { }
	attributes_t(const cudaFuncAttributes& a) : cudaFuncAttributes(a) { }
};

/// The dynamic shared memory a block of this kernel may use on this device: the opt-in per-block
/// maximum less the kernel's static shared memory.  (The reference looks a per-architecture
/// default up in a table which ends at compute capability 6.2,
/// external/.../device_properties.cpp:27-40.)
inline shared_memory_size_t maximum_dynamic_shared_memory_per_block(
	const attributes_t& attributes, const device::properties_t& properties)
{
	::std::size_t per_block = properties.sharedMemPerBlockOptin ? properties.sharedMemPerBlockOptin : properties.sharedMemPerBlock;
	return (shared_memory_size_t)(per_block > attributes.sharedSizeBytes ? per_block - attributes.sharedSizeBytes : 0);
}

} // namespace device_function

/// A handle on a __global__ function which lives inside libgiddy_b200.so.  Attribute queries and
/// launches go through the library, because the function is registered with the library's CUDA
/// runtime, not necessarily the caller's.
class device_function_t {
public:
	device_function_t(const void* f) : ptr_(f) { }
	const void* ptr() const { return ptr_; }
	operator const void*() const { return ptr_; }

	device_function::attributes_t attributes() const
	{
		giddy_function_attributes_t a;
		detail::throw_if_error(giddy_function_attributes(ptr_, &a));
		device_function::attributes_t r;
		r.maxThreadsPerBlock = a.max_threads_per_block;
		r.numRegs = a.num_regs;
		r.sharedSizeBytes = a.shared_size_bytes;
		r.constSizeBytes = a.const_size_bytes;
		r.localSizeBytes = a.local_size_bytes;
		r.maxDynamicSharedSizeBytes = a.max_dynamic_shared_size_bytes;
		r.ptxVersion = a.ptx_version;
		r.binaryVersion = a.binary_version;
		return r;
	}
protected:
	const void* ptr_;
};

/// Launch a library kernel with a caller-assembled parameter pack (reference: cuda::enqueue_launch,
/// external/.../kernel_launch.cuh:90-114).  Parameter types must match the kernel's exactly.
template <typename... KernelParameters>
inline void enqueue_launch(
	const device_function_t& kernel_function, stream::id_t stream_id,
	const launch_configuration_t& launch_config, KernelParameters... parameters)
{
	void* argument_pointers[] = { const_cast<void*>(static_cast<const void*>(&parameters))... , nullptr };
	giddy_launch_config_t config = to_c(launch_config);
	detail::throw_if_error(giddy_launch_kernel(kernel_function.ptr(), stream_id, &config, argument_pointers));
}

} // namespace cuda

#endif // GIDDY_TYPES_HPP_
