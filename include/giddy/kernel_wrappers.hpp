// giddy/kernel_wrappers.hpp -- the per-scheme kernel wrappers and their launch-configuration
// parameters: eleven class templates cuda::kernels::<scheme>::kernel_t<...> with, each,
//   * a typed resolve_launch_configuration(props, attrs, scheme scalars..., limits)       ("mode ii")
//   * the two type-erased overrides taking name -> any maps                               ("mode iii")
//   * get_device_function(), the address of the sm_100a kernel inside libgiddy_b200.so
//   * key(), the factory key -- the reference's demangled class name
// and eleven cuda::kernels::<scheme>::launch_config_resolution_params_t<...> for "mode i".
//
// Class names, template parameter lists, argument-map keys and their exact static types follow the
// reference's src/kernel_wrappers/**/*.cu (cited per wrapper).  The launch-configuration *policy*
// inside each params class is new: it describes how the B200 kernels want their grids built.
//
// Everything below only calls the C ABI (giddy_b200.h); no CUDA compiler is needed to use it.
#pragma once
#ifndef GIDDY_KERNEL_WRAPPERS_HPP_
#define GIDDY_KERNEL_WRAPPERS_HPP_

#include "registered_wrapper.hpp"
#include "resolve_launch_configuration.hpp"
#include "model_functors.hpp"
#include "kernel_traits.h"

#include <string>

namespace cuda {
namespace kernels {

namespace detail {

template <typename T> struct type_name_of;
template <> struct type_name_of<::std::int8_t>   { static const char* get() { return "signed char"; } };
template <> struct type_name_of<::std::int16_t>  { static const char* get() { return "short"; } };
template <> struct type_name_of<::std::int32_t>  { static const char* get() { return "int"; } };
template <> struct type_name_of<::std::int64_t>  { static const char* get() { return "long"; } };
template <> struct type_name_of<::std::uint8_t>  { static const char* get() { return "unsigned char"; } };
template <> struct type_name_of<::std::uint16_t> { static const char* get() { return "unsigned short"; } };
template <> struct type_name_of<::std::uint32_t> { static const char* get() { return "unsigned int"; } };
template <> struct type_name_of<::std::uint64_t> { static const char* get() { return "unsigned long"; } };

template <unsigned I, typename T>
struct type_name_of<functors::unary::parametric_model::linear<I, T>> {
	static ::std::string get() { return "cuda::functors::unary::parametric_model::linear<" + ::std::to_string(I) + "u, " + type_name_of<T>::get() + ">"; }
};
template <unsigned I, typename T>
struct type_name_of<functors::unary::parametric_model::constant<I, T>> {
	static ::std::string get() { return "cuda::functors::unary::parametric_model::constant<" + ::std::to_string(I) + "u, " + type_name_of<T>::get() + ">"; }
};

inline ::std::string u(unsigned x) { return ::std::to_string(x) + "u"; }

inline giddy_launch_config_t checked(const launch_configuration_t& c) { return to_c(c); }

/// any_cast of a required argument: std::out_of_range when missing, bad_any_cast on a type mismatch
template <typename T>
inline T required(const kernel_t::arguments_type& arguments, const char* name) { return ::cuda::any_cast<T>(arguments.at(name)); }

template <typename T>
inline optional<T> optional_argument(const kernel_t::arguments_type& arguments, const char* name) { return ::cuda::any_cast<T>(maybe_at(arguments, name)); }

/// The kernel's attributes: those the caller already has, or else ask the runtime (one driver call)
inline device_function::attributes_t attributes_or_query(const optional<device_function::attributes_t>& known, const void* kernel)
{
	return known ? known.value() : device_function_t(kernel).attributes();
}

/// The configuration shared by the warp-per-unit kernels: one length unit per warp, a fixed number of warps per block
inline void configure_warp_per_unit(launch_config_resolution_params_t& p, ::std::size_t num_units, unsigned warps_per_block)
{
	p.grid_construction_resolution  = launch_config_resolution_params_t::warp;
	p.serialization_option          = launch_config_resolution_params_t::none;
	p.length                        = num_units;
	p.quanta.threads_in_block       = warp_size;
	p.upper_limits.warps_in_block   = warps_per_block;
}

} // namespace detail

#define GIDDY_WRAPPER_BOILERPLATE \
	using parent          = cuda::registered::kernel_t; \
	using arguments_type  = parent::arguments_type; \
	using parent::resolve_launch_configuration; \
	static void registerInSubclassFactory() { parent::template registerInSubclassFactory<kernel_t>(key()); } \
	launch_configuration_t resolve_launch_configuration( \
		device::properties_t           device_properties, \
		device_function::attributes_t  kernel_function_attributes, \
		arguments_type                 extra_arguments, \
		launch_configuration_limits_t  limits = { nullopt, nullopt, nullopt }) const override final; \
	/* Note: the stream is assumed to be on the current device */ \
	void enqueue_launch( \
		stream::id_t                   stream, \
		const launch_configuration_t&  launch_config, \
		arguments_type                 arguments) const override final; \
	const device_function_t get_device_function() const override final

namespace decompression {

// =================================================================================================
// Frame of Reference -- reference src/kernel_wrappers/decompression/frame_of_reference.cu:22-146
// =================================================================================================
namespace frame_of_reference {

template <unsigned IndexSize, typename Uncompressed, typename Compressed, typename UnaryModelFunction>
inline const void* kernel_address()
{
	return giddy_b200_frame_of_reference_kernel(IndexSize, sizeof(Uncompressed), sizeof(Compressed),
		(giddy_model_t) UnaryModelFunction::kind, UnaryModelFunction::index_size);
}

template <unsigned IndexSize, typename Uncompressed, typename Compressed, typename UnaryModelFunction>
class launch_config_resolution_params_t final : public kernels::launch_config_resolution_params_t {
public:
	using parent = kernels::launch_config_resolution_params_t;
	launch_config_resolution_params_t(device::properties_t device_properties_, ::std::size_t data_length, ::std::size_t segment_length,
		optional<device_function::attributes_t> known_attributes = nullopt) :
		parent(device_properties_, detail::attributes_or_query(known_attributes, kernel_address<IndexSize, Uncompressed, Compressed, UnaryModelFunction>()))
	{
		// the same preconditions as the reference, frame_of_reference.cuh:99-104
		if (segment_length == 0) { throw ::std::invalid_argument("The segment length must be a positive number"); }
		if (segment_length % sizeof(native_word_t) != 0) {
			throw ::std::invalid_argument("The segment length must be aligned with the GPU's native word size");
		}
		grid_construction_resolution   = thread;
		length                         = data_length;
		serialization_option           = fixed_factor;
		default_serialization_factor   = (16 / sizeof(Uncompressed)) * GIDDY_ELEMENTWISE_UNROLL;
		quanta.threads_in_block        = warp_size;
		upper_limits.threads_in_block  = GIDDY_ELEMENTWISE_BLOCK_THREADS;
	}
};

template <unsigned IndexSize, typename Uncompressed, typename Compressed, typename UnaryModelFunction>
class kernel_t : public cuda::registered::kernel_t {
public:
	GIDDY_WRAPPER_BOILERPLATE;

	using size_type               = size_type_by_index_size<IndexSize>;
	using model_coefficients_type = typename UnaryModelFunction::coefficients_type;

	static ::std::string key()
	{
		return "decompression::frame_of_reference::kernel_t<" + detail::u(IndexSize) + ", " +
			detail::type_name_of<Uncompressed>::get() + ", " + detail::type_name_of<Compressed>::get() + ", " +
			detail::type_name_of<UnaryModelFunction>::get() + " >";
	}

	launch_configuration_t resolve_launch_configuration(
		device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
		::std::size_t length, ::std::size_t segment_length,
		launch_configuration_limits_t limits = { nullopt, nullopt, nullopt }) const
	{
		launch_config_resolution_params_t<IndexSize, Uncompressed, Compressed, UnaryModelFunction> params(device_properties, length, segment_length, optional<device_function::attributes_t>(kernel_function_attributes));
		return kernels::resolve_launch_configuration(params, limits);
	}

	void enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config,
		Uncompressed* decompressed, const Compressed* compressed_input, const model_coefficients_type* segment_model_coefficients,
		uint_t<IndexSize> length, uint_t<IndexSize> segment_length) const
	{
		auto config = to_c(launch_config);
		cuda::detail::throw_if_error(giddy_b200_frame_of_reference_launch(stream, &config,
			IndexSize, sizeof(Uncompressed), sizeof(Compressed), (giddy_model_t) UnaryModelFunction::kind, UnaryModelFunction::index_size,
			decompressed, compressed_input, segment_model_coefficients, length, segment_length));
	}
};

template <unsigned I, typename U, typename C, typename M>
launch_configuration_t kernel_t<I, U, C, M>::resolve_launch_configuration(
	device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
	arguments_type extra_arguments, launch_configuration_limits_t limits) const
{
	auto length         = detail::required<::std::size_t>(extra_arguments, "length");
	auto segment_length = detail::required<::std::size_t>(extra_arguments, "segment_length");
	return resolve_launch_configuration(device_properties, kernel_function_attributes, length, segment_length, limits);
}

template <unsigned I, typename U, typename C, typename M>
void kernel_t<I, U, C, M>::enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config, arguments_type arguments) const
{
	using index_type = uint_t<I>;
	auto decompressed               = detail::required<U*>(arguments, "decompressed");
	auto compressed_input           = detail::required<const C*>(arguments, "compressed_input");
	auto segment_model_coefficients = detail::required<const model_coefficients_type*>(arguments, "segment_model_coefficients");
	auto length                     = detail::required<index_type>(arguments, "length");
	auto segment_length             = detail::required<index_type>(arguments, "segment_length");
	enqueue_launch(stream, launch_config, decompressed, compressed_input, segment_model_coefficients, length, segment_length);
}

template <unsigned I, typename U, typename C, typename M>
const device_function_t kernel_t<I, U, C, M>::get_device_function() const { return kernel_address<I, U, C, M>(); }

} // namespace frame_of_reference

// =================================================================================================
// Model -- reference src/kernel_wrappers/decompression/model.cu:13-98
// =================================================================================================
namespace model {

template <unsigned IndexSize, typename Uncompressed, typename UnaryModelFunction>
inline const void* kernel_address()
{
	return giddy_b200_model_kernel(IndexSize, sizeof(Uncompressed), (giddy_model_t) UnaryModelFunction::kind);
}

template <unsigned IndexSize, typename Uncompressed, typename UnaryModelFunction>
class launch_config_resolution_params_t final : public kernels::launch_config_resolution_params_t {
public:
	using parent = kernels::launch_config_resolution_params_t;
	launch_config_resolution_params_t(device::properties_t device_properties_, ::std::size_t length_,
		optional<device_function::attributes_t> known_attributes = nullopt) :
		parent(device_properties_, detail::attributes_or_query(known_attributes, kernel_address<IndexSize, Uncompressed, UnaryModelFunction>()))
	{
		grid_construction_resolution   = thread;
		length                         = length_;
		serialization_option           = fixed_factor;
		default_serialization_factor   = (16 / sizeof(Uncompressed)) * GIDDY_ELEMENTWISE_UNROLL;
		quanta.threads_in_block        = warp_size;
		upper_limits.threads_in_block  = GIDDY_ELEMENTWISE_BLOCK_THREADS;
	}
};

template <unsigned IndexSize, typename Uncompressed, typename UnaryModelFunction>
class kernel_t : public cuda::registered::kernel_t {
	static_assert(::std::is_same<Uncompressed, typename UnaryModelFunction::coefficient_type>::value,
		"the model's coefficient type must be the uncompressed type");
public:
	GIDDY_WRAPPER_BOILERPLATE;
	using model_coefficients_type = typename UnaryModelFunction::coefficients_type;

	static ::std::string key()
	{
		return "decompression::model::kernel_t<" + detail::u(IndexSize) + ", " + detail::type_name_of<Uncompressed>::get() + ", " +
			detail::type_name_of<UnaryModelFunction>::get() + " >";
	}

	launch_configuration_t resolve_launch_configuration(
		device::properties_t device_properties, device_function::attributes_t kernel_function_attributes, ::std::size_t length,
		launch_configuration_limits_t limits = { nullopt, nullopt, nullopt }) const
	{
		launch_config_resolution_params_t<IndexSize, Uncompressed, UnaryModelFunction> params(device_properties, length, optional<device_function::attributes_t>(kernel_function_attributes));
		return kernels::resolve_launch_configuration(params, limits);
	}

	void enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config,
		Uncompressed* decompressed, model_coefficients_type model_coefficients, uint_t<IndexSize> length) const
	{
		auto config = to_c(launch_config);
		cuda::detail::throw_if_error(giddy_b200_model_launch(stream, &config, IndexSize, sizeof(Uncompressed),
			(giddy_model_t) UnaryModelFunction::kind, decompressed, model_coefficients.data(), length));
	}
};

template <unsigned I, typename U, typename M>
launch_configuration_t kernel_t<I, U, M>::resolve_launch_configuration(
	device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
	arguments_type extra_arguments, launch_configuration_limits_t limits) const
{
	auto length = detail::required<::std::size_t>(extra_arguments, "length");
	return resolve_launch_configuration(device_properties, kernel_function_attributes, length, limits);
}

template <unsigned I, typename U, typename M>
void kernel_t<I, U, M>::enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config, arguments_type arguments) const
{
	auto length             = detail::required<uint_t<I>>(arguments, "length");
	auto decompressed       = detail::required<U*>(arguments, "decompressed");
	auto model_coefficients = detail::required<model_coefficients_type>(arguments, "model_coefficients");
	enqueue_launch(stream, launch_config, decompressed, model_coefficients, length);
}

template <unsigned I, typename U, typename M>
const device_function_t kernel_t<I, U, M>::get_device_function() const { return kernel_address<I, U, M>(); }

} // namespace model

// =================================================================================================
// Run-Length Encoding -- reference src/kernel_wrappers/decompression/run_length_encoding.cu:13-128
// =================================================================================================
namespace run_length_encoding {

template <unsigned UncompressedIndexSize, unsigned UncompressedSize, unsigned RunLengthSize, unsigned RunIndexSize = UncompressedIndexSize>
class launch_config_resolution_params_t final : public kernels::launch_config_resolution_params_t {
public:
	using parent = kernels::launch_config_resolution_params_t;
	launch_config_resolution_params_t(device::properties_t device_properties_, ::std::size_t uncompressed_length, ::std::size_t segment_length,
		optional<shared_memory_size_t> dynamic_shared_mem_limit = nullopt,
		optional<device_function::attributes_t> known_attributes = nullopt) :
		parent(device_properties_, detail::attributes_or_query(known_attributes, giddy_b200_run_length_encoding_kernel(UncompressedIndexSize, UncompressedSize, RunLengthSize, RunIndexSize)), dynamic_shared_mem_limit)
	{
		if (segment_length == 0) { throw ::std::invalid_argument("The segment length must be a positive number"); }
		detail::configure_warp_per_unit(*this, util::div_rounding_up(uncompressed_length, segment_length), GIDDY_RUNS_WARPS_PER_BLOCK);
		dynamic_shared_memory_requirement.per_length_unit = GIDDY_RUNS_SHARED_BYTES_PER_WARP(UncompressedSize) / warp_size;
	}
};

template <unsigned UncompressedIndexSize, unsigned UncompressedSize, unsigned RunLengthSize, unsigned RunIndexSize = UncompressedIndexSize>
class kernel_t : public cuda::registered::kernel_t {
public:
	GIDDY_WRAPPER_BOILERPLATE;

	using uncompressed_type    = uint_t<UncompressedSize>;
	using run_length_type      = uint_t<RunLengthSize>;
	using run_index_type       = uint_t<RunIndexSize>;
	using size_type            = size_type_by_index_size<UncompressedIndexSize>;
	using run_index_size_type  = size_type_by_index_size<RunIndexSize>;

	static ::std::string key()
	{
		return "decompression::run_length_encoding::kernel_t<" + detail::u(UncompressedIndexSize) + ", " + detail::u(UncompressedSize) + ", " +
			detail::u(RunLengthSize) + ", " + detail::u(RunIndexSize) + ">";
	}

	launch_configuration_t resolve_launch_configuration(
		device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
		::std::size_t uncompressed_length, ::std::size_t segment_length,
		launch_configuration_limits_t limits = { nullopt, nullopt, nullopt }) const
	{
		launch_config_resolution_params_t<UncompressedIndexSize, UncompressedSize, RunLengthSize, RunIndexSize> params(
			device_properties, uncompressed_length, segment_length, limits.dynamic_shared_memory, optional<device_function::attributes_t>(kernel_function_attributes));
		return kernels::resolve_launch_configuration(params, limits);
	}

	void enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config,
		uncompressed_type* decompressed, const uncompressed_type* run_data, const run_length_type* run_lengths,
		const run_index_type* position_anchors, const run_length_type* intra_run_anchor_offsets,
		size_type segment_length, size_type num_segments, run_index_size_type num_element_runs, size_type uncompressed_length) const
	{
		auto config = to_c(launch_config);
		cuda::detail::throw_if_error(giddy_b200_run_length_encoding_launch(stream, &config,
			UncompressedIndexSize, UncompressedSize, RunLengthSize, RunIndexSize,
			decompressed, run_data, run_lengths, position_anchors, intra_run_anchor_offsets,
			segment_length, num_segments, num_element_runs, uncompressed_length));
	}
};

template <unsigned UI, unsigned U, unsigned RL, unsigned RI>
launch_configuration_t kernel_t<UI, U, RL, RI>::resolve_launch_configuration(
	device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
	arguments_type extra_arguments, launch_configuration_limits_t limits) const
{
	auto uncompressed_length = detail::required<::std::size_t>(extra_arguments, "length");
	auto segment_length      = detail::required<::std::size_t>(extra_arguments, "segment_length");
	return resolve_launch_configuration(device_properties, kernel_function_attributes, uncompressed_length, segment_length, limits);
}

template <unsigned UI, unsigned U, unsigned RL, unsigned RI>
void kernel_t<UI, U, RL, RI>::enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config, arguments_type arguments) const
{
	auto decompressed             = detail::required<uncompressed_type*>(arguments, "decompressed");
	auto run_data                 = detail::required<const uncompressed_type*>(arguments, "run_data");
	auto run_lengths              = detail::required<const run_length_type*>(arguments, "run_lengths");
	auto position_anchors         = detail::required<const run_index_type*>(arguments, "position_anchors");
	auto intra_run_anchor_offsets = detail::required<const run_length_type*>(arguments, "intra_run_anchor_offsets");
	auto segment_length           = detail::required<size_type>(arguments, "segment_length");
	auto num_segments             = detail::required<size_type>(arguments, "num_segments");
	auto num_element_runs         = detail::required<run_index_size_type>(arguments, "num_element_runs");
	auto uncompressed_length      = detail::required<size_type>(arguments, "length");
	enqueue_launch(stream, launch_config, decompressed, run_data, run_lengths, position_anchors, intra_run_anchor_offsets,
		segment_length, num_segments, num_element_runs, uncompressed_length);
}

template <unsigned UI, unsigned U, unsigned RL, unsigned RI>
const device_function_t kernel_t<UI, U, RL, RI>::get_device_function() const
{
	return giddy_b200_run_length_encoding_kernel(UI, U, RL, RI);
}

} // namespace run_length_encoding

// =================================================================================================
// Run-Position Encoding -- reference src/kernel_wrappers/decompression/run_position_encoding.cu:13-136
// =================================================================================================
namespace run_position_encoding {

template <unsigned UncompressedIndexSize, unsigned UncompressedSize, unsigned PositionOffsetSize, bool PositionsAreRelative, unsigned RunIndexSize = UncompressedIndexSize>
class launch_config_resolution_params_t final : public kernels::launch_config_resolution_params_t {
public:
	using parent = kernels::launch_config_resolution_params_t;
	launch_config_resolution_params_t(device::properties_t device_properties_, ::std::size_t uncompressed_length, ::std::size_t position_anchoring_period,
		optional<shared_memory_size_t> dynamic_shared_mem_limit = nullopt,
		optional<device_function::attributes_t> known_attributes = nullopt) :
		parent(device_properties_, detail::attributes_or_query(known_attributes, giddy_b200_run_position_encoding_kernel(
			UncompressedIndexSize, UncompressedSize, PositionOffsetSize, PositionsAreRelative, RunIndexSize)), dynamic_shared_mem_limit)
	{
		if (position_anchoring_period == 0) { throw ::std::invalid_argument("The segment length must be a positive number"); }
		detail::configure_warp_per_unit(*this, util::div_rounding_up(uncompressed_length, position_anchoring_period), GIDDY_RUNS_WARPS_PER_BLOCK);
		dynamic_shared_memory_requirement.per_length_unit = GIDDY_RUNS_SHARED_BYTES_PER_WARP(UncompressedSize) / warp_size;
	}
};

template <unsigned UncompressedIndexSize, unsigned UncompressedSize, unsigned PositionOffsetSize, bool PositionsAreRelative, unsigned RunIndexSize = UncompressedIndexSize>
class kernel_t : public cuda::registered::kernel_t {
	static_assert(PositionsAreRelative || PositionOffsetSize >= UncompressedIndexSize,
		"absolute run positions must be able to cover the entire potential range of data");
public:
	GIDDY_WRAPPER_BOILERPLATE;

	using uncompressed_type             = uint_t<UncompressedSize>;
	using position_offset_type          = uint_t<PositionOffsetSize>;
	using run_index_type                = uint_t<RunIndexSize>;
	using uncompressed_index_size_type  = size_type_by_index_size<UncompressedIndexSize>;
	using run_index_size_type           = size_type_by_index_size<RunIndexSize>;
	using position_offset_size_type     = size_type_by_index_size<PositionOffsetSize>;

	static ::std::string key()
	{
		return "decompression::run_position_encoding::kernel_t<" + detail::u(UncompressedIndexSize) + ", " + detail::u(UncompressedSize) + ", " +
			detail::u(PositionOffsetSize) + ", " + (PositionsAreRelative ? "true" : "false") + ", " + detail::u(RunIndexSize) + ">";
	}

	launch_configuration_t resolve_launch_configuration(
		device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
		::std::size_t uncompressed_length, ::std::size_t segment_length,
		launch_configuration_limits_t limits = { nullopt, nullopt, nullopt }) const
	{
		launch_config_resolution_params_t<UncompressedIndexSize, UncompressedSize, PositionOffsetSize, PositionsAreRelative, RunIndexSize> params(
			device_properties, uncompressed_length, segment_length, limits.dynamic_shared_memory, optional<device_function::attributes_t>(kernel_function_attributes));
		return kernels::resolve_launch_configuration(params, limits);
	}

	void enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config,
		uncompressed_type* decompressed, const uncompressed_type* run_data, const position_offset_type* run_start_positions,
		const run_index_type* position_anchors, position_offset_size_type segment_length,
		uncompressed_index_size_type num_anchors, run_index_size_type num_element_runs, uncompressed_index_size_type uncompressed_length) const
	{
		auto config = to_c(launch_config);
		cuda::detail::throw_if_error(giddy_b200_run_position_encoding_launch(stream, &config,
			UncompressedIndexSize, UncompressedSize, PositionOffsetSize, PositionsAreRelative, RunIndexSize,
			decompressed, run_data, run_start_positions, position_anchors, segment_length, num_anchors, num_element_runs, uncompressed_length));
	}
};

template <unsigned UI, unsigned U, unsigned PO, bool Rel, unsigned RI>
launch_configuration_t kernel_t<UI, U, PO, Rel, RI>::resolve_launch_configuration(
	device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
	arguments_type extra_arguments, launch_configuration_limits_t limits) const
{
	auto uncompressed_length = detail::required<::std::size_t>(extra_arguments, "length");
	auto segment_length      = detail::required<::std::size_t>(extra_arguments, "segment_length");
	return resolve_launch_configuration(device_properties, kernel_function_attributes, uncompressed_length, segment_length, limits);
}

template <unsigned UI, unsigned U, unsigned PO, bool Rel, unsigned RI>
void kernel_t<UI, U, PO, Rel, RI>::enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config, arguments_type arguments) const
{
	auto decompressed        = detail::required<uncompressed_type*>(arguments, "decompressed");
	auto run_data            = detail::required<const uncompressed_type*>(arguments, "run_data");
	auto run_start_positions = detail::required<const position_offset_type*>(arguments, "run_start_positions");
	auto position_anchors    = detail::required<const run_index_type*>(arguments, "position_anchors");
	auto segment_length      = detail::required<position_offset_size_type>(arguments, "segment_length");
	auto num_anchors         = detail::required<uncompressed_index_size_type>(arguments, "num_anchors");
	auto num_element_runs    = detail::required<run_index_size_type>(arguments, "num_element_runs");
	auto uncompressed_length = detail::required<uncompressed_index_size_type>(arguments, "length");
	enqueue_launch(stream, launch_config, decompressed, run_data, run_start_positions, position_anchors,
		segment_length, num_anchors, num_element_runs, uncompressed_length);
}

template <unsigned UI, unsigned U, unsigned PO, bool Rel, unsigned RI>
const device_function_t kernel_t<UI, U, PO, Rel, RI>::get_device_function() const
{
	return giddy_b200_run_position_encoding_kernel(UI, U, PO, Rel, RI);
}

} // namespace run_position_encoding

// =================================================================================================
// Delta -- reference src/kernel_wrappers/decompression/delta.cu:13-108
// =================================================================================================
namespace delta {

template <unsigned IndexSize, unsigned UncompressedSize, unsigned CompressedSize>
class launch_config_resolution_params_t final : public kernels::launch_config_resolution_params_t {
public:
	using parent = kernels::launch_config_resolution_params_t;
	launch_config_resolution_params_t(device::properties_t device_properties_, ::std::size_t data_length, ::std::size_t segment_length,
		optional<shared_memory_size_t> = nullopt,
		optional<device_function::attributes_t> known_attributes = nullopt) :
		parent(device_properties_, detail::attributes_or_query(known_attributes, giddy_b200_delta_kernel(IndexSize, UncompressedSize, CompressedSize)))
	{
		if (segment_length == 0) { throw ::std::invalid_argument("The segment length must be a positive number"); }
		if (GIDDY_DELTA_USES_LOOK_BACK(data_length, segment_length, UncompressedSize)) {
			// few, long segments: tiles scanned with decoupled look-back by persistent CTAs fed from a ticket counter;
			// any grid is correct, one resident wave is what is wanted
			const ::std::size_t tile = GIDDY_DELTA_TILE_ELEMENTS(UncompressedSize, GIDDY_DELTA_LOOK_BACK_BLOCK_THREADS);
			grid_construction_resolution  = block;
			serialization_option          = auto_maximized;
			length                        = util::div_rounding_up(data_length, tile);
			quanta.threads_in_block       = warp_size;
			upper_limits.threads_in_block = GIDDY_DELTA_LOOK_BACK_BLOCK_THREADS;
			block_resolution_constraints.max_threads_per_block = GIDDY_DELTA_LOOK_BACK_BLOCK_THREADS;
			keep_gpu_busy_factor          = GIDDY_DELTA_LOOK_BACK_BLOCKS_PER_MULTIPROCESSOR * (GIDDY_DELTA_LOOK_BACK_BLOCK_THREADS / warp_size) / 4;
			return;
		}
		detail::configure_warp_per_unit(*this, util::div_rounding_up(data_length, segment_length), GIDDY_SEGMENT_WARPS_PER_BLOCK);
	}
};

template <unsigned IndexSize, unsigned UncompressedSize, unsigned CompressedSize>
class kernel_t : public cuda::registered::kernel_t {
public:
	GIDDY_WRAPPER_BOILERPLATE;
	using uncompressed_type = uint_t<UncompressedSize>;
	using compressed_type   = uint_t<CompressedSize>;

	static ::std::string key()
	{
		return "decompression::delta::kernel_t<" + detail::u(IndexSize) + ", " + detail::u(UncompressedSize) + ", " + detail::u(CompressedSize) + ">";
	}

	launch_configuration_t resolve_launch_configuration(
		device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
		::std::size_t length, ::std::size_t segment_length,
		launch_configuration_limits_t limits = { nullopt, nullopt, nullopt }) const
	{
		launch_config_resolution_params_t<IndexSize, UncompressedSize, CompressedSize> params(device_properties, length, segment_length, nullopt, optional<device_function::attributes_t>(kernel_function_attributes));
		return kernels::resolve_launch_configuration(params, limits);
	}

	void enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config,
		uncompressed_type* decompressed, const compressed_type* compressed_input, const uncompressed_type* anchor_values,
		uint_t<IndexSize> length, uint_t<IndexSize> segment_length) const
	{
		auto config = to_c(launch_config);
		cuda::detail::throw_if_error(giddy_b200_delta_launch(stream, &config, IndexSize, UncompressedSize, CompressedSize,
			decompressed, compressed_input, anchor_values, length, segment_length));
	}
};

template <unsigned I, unsigned U, unsigned C>
launch_configuration_t kernel_t<I, U, C>::resolve_launch_configuration(
	device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
	arguments_type extra_arguments, launch_configuration_limits_t limits) const
{
	auto length         = detail::required<::std::size_t>(extra_arguments, "length");
	auto segment_length = detail::required<::std::size_t>(extra_arguments, "segment_length");
	return resolve_launch_configuration(device_properties, kernel_function_attributes, length, segment_length, limits);
}

template <unsigned I, unsigned U, unsigned C>
void kernel_t<I, U, C>::enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config, arguments_type arguments) const
{
	auto decompressed     = detail::required<uncompressed_type*>(arguments, "decompressed");
	auto compressed_input = detail::required<const compressed_type*>(arguments, "compressed_input");
	auto anchor_values    = detail::required<const uncompressed_type*>(arguments, "anchor_values");
	auto length           = detail::required<uint_t<I>>(arguments, "length");
	auto segment_length   = detail::required<uint_t<I>>(arguments, "segment_length");
	enqueue_launch(stream, launch_config, decompressed, compressed_input, anchor_values, length, segment_length);
}

template <unsigned I, unsigned U, unsigned C>
const device_function_t kernel_t<I, U, C>::get_device_function() const { return giddy_b200_delta_kernel(I, U, C); }

} // namespace delta

// =================================================================================================
// Dictionary -- reference src/kernel_wrappers/decompression/dictionary.cu:13-141 (params: gather.cuh:166-223)
// =================================================================================================
namespace dictionary {

/// Dictionaries up to this size are, by default, staged in shared memory by every CTA
enum : unsigned { max_default_cached_dictionary_bytes = 32 * 1024 };
/// ... and larger ones by a persistent grid, if at least this much shared memory is to be had for them
enum : unsigned { min_large_cached_dictionary_bytes = 96 * 1024 };

template <unsigned IndexSize, unsigned UncompressedSize, unsigned DictionaryIndexSize>
class launch_config_resolution_params_t final : public kernels::launch_config_resolution_params_t {
public:
	using parent = kernels::launch_config_resolution_params_t;
	launch_config_resolution_params_t(device::properties_t device_properties_, ::std::size_t input_data_length /* dictionary entries */,
		::std::size_t num_indices /* data length */, optional<bool> cache_input_data_in_shared_memory_override = nullopt,
		optional<shared_memory_size_t> dynamic_shared_mem_limit = nullopt,
		optional<device_function::attributes_t> known_attributes = nullopt) :
		parent(device_properties_, detail::attributes_or_query(known_attributes, giddy_b200_dictionary_kernel(IndexSize, UncompressedSize, DictionaryIndexSize)), dynamic_shared_mem_limit)
	{
		const ::std::size_t dictionary_bytes = input_data_length * UncompressedSize;
		const ::std::size_t output_bytes = num_indices * UncompressedSize;
		// The reference never caches by default ("seems to mostly hurt us", gather.cuh:202-205 -- its caching
		// path is commented out).  Here
		//  * a small dictionary is read once per CTA and then served from shared memory, provided the column is long
		//    enough to amortise that;
		//  * a large one (random lookups into it run at about half of HBM speed through L1/L2) is worth a persistent
		//    grid of one big CTA per multiprocessor, each holding the dictionary -- or as much of its head as fits -- in
		//    up to 227 KiB of shared memory, once the column is ~1000 times the dictionary's size.
		const bool small_worthwhile = dictionary_bytes <= max_default_cached_dictionary_bytes &&
			dictionary_bytes <= available_dynamic_shared_memory_per_block && output_bytes >= 64 * dictionary_bytes;
		const bool large_worthwhile = dictionary_bytes > max_default_cached_dictionary_bytes &&
			available_dynamic_shared_memory_per_block >= min_large_cached_dictionary_bytes && output_bytes / 1024 >= dictionary_bytes;
		const bool cache = cache_input_data_in_shared_memory_override.value_or(small_worthwhile || large_worthwhile);
		grid_construction_resolution   = thread;
		length                         = num_indices;
		quanta.threads_in_block        = warp_size;
		if (cache && dictionary_bytes > max_default_cached_dictionary_bytes) {
			const ::std::size_t room = (available_dynamic_shared_memory_per_block / 16) * 16;
			dynamic_shared_memory_requirement.per_block = (shared_memory_size_t) ::std::min(dictionary_bytes, room);
			serialization_option       = auto_maximized;
			keep_gpu_busy_factor       = 8; // x 4 schedulers / 32 warps: one CTA of 1024 threads per multiprocessor
			return;
		}
		serialization_option           = runtime_specified_factor;
		default_serialization_factor   = (16 / UncompressedSize) * GIDDY_ELEMENTWISE_UNROLL;
		upper_limits.threads_in_block  = GIDDY_ELEMENTWISE_BLOCK_THREADS;
		if (cache) { dynamic_shared_memory_requirement.per_block = (shared_memory_size_t) dictionary_bytes; }
	}
};

template <unsigned IndexSize, unsigned UncompressedSize, unsigned DictionaryIndexSize>
class kernel_t : public cuda::registered::kernel_t {
public:
	GIDDY_WRAPPER_BOILERPLATE;
	using dictionary_index_type = uint_t<DictionaryIndexSize>;
	using dictionary_size_type  = size_type_by_index_size<DictionaryIndexSize>;
	using uncompressed_type     = uint_t<UncompressedSize>;
	using size_type             = size_type_by_index_size<IndexSize>;

	static ::std::string key()
	{
		return "decompression::dictionary::kernel_t<" + detail::u(IndexSize) + ", " + detail::u(UncompressedSize) + ", " + detail::u(DictionaryIndexSize) + ">";
	}

	launch_configuration_t resolve_launch_configuration(
		device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
		::std::size_t num_dictionary_entries, ::std::size_t data_length,
		optional<bool> cache_input_data_in_shared_mem = nullopt, optional<serialization_factor_t> serialization_factor = nullopt,
		launch_configuration_limits_t limits = { nullopt, nullopt, nullopt }) const
	{
		launch_config_resolution_params_t<IndexSize, UncompressedSize, DictionaryIndexSize> params(
			device_properties, num_dictionary_entries, data_length, cache_input_data_in_shared_mem, limits.dynamic_shared_memory, optional<device_function::attributes_t>(kernel_function_attributes));
		return kernels::resolve_launch_configuration(params, limits, serialization_factor);
	}

	void enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config,
		uncompressed_type* decompressed, const dictionary_index_type* compressed_input, const uncompressed_type* dictionary_entries,
		size_type length, dictionary_size_type num_dictionary_entries, bool cache_dictionary_in_shared_mem, serialization_factor_t serialization_factor) const
	{
		auto config = to_c(launch_config);
		cuda::detail::throw_if_error(giddy_b200_dictionary_launch(stream, &config, IndexSize, UncompressedSize, DictionaryIndexSize,
			decompressed, compressed_input, dictionary_entries, length, num_dictionary_entries, cache_dictionary_in_shared_mem, serialization_factor));
	}
};

template <unsigned I, unsigned U, unsigned DI>
launch_configuration_t kernel_t<I, U, DI>::resolve_launch_configuration(
	device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
	arguments_type extra_arguments, launch_configuration_limits_t limits) const
{
	auto num_dictionary_entries         = detail::required<::std::size_t>(extra_arguments, "num_dictionary_entries");
	auto data_length                    = detail::required<::std::size_t>(extra_arguments, "data_length");
	auto cache_input_data_in_shared_mem = detail::optional_argument<bool>(extra_arguments, "cache_input_data_in_shared_mem");
	auto serialization_factor           = detail::optional_argument<serialization_factor_t>(extra_arguments, "serialization_factor");
	return resolve_launch_configuration(device_properties, kernel_function_attributes,
		num_dictionary_entries, data_length, cache_input_data_in_shared_mem, serialization_factor, limits);
}

template <unsigned I, unsigned U, unsigned DI>
void kernel_t<I, U, DI>::enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config, arguments_type arguments) const
{
	auto caching = detail::optional_argument<bool>(arguments, "cache_dictionary_in_shared_mem").value_or(launch_config.dynamic_shared_memory_size > 0);
	auto serialization_factor = detail::optional_argument<serialization_factor_t>(arguments, "serialization_factor").value_or(
		(serialization_factor_t)((16 / U) * GIDDY_ELEMENTWISE_UNROLL));
	auto decompressed           = detail::required<uncompressed_type*>(arguments, "decompressed");
	auto compressed_input       = detail::required<const dictionary_index_type*>(arguments, "compressed_input");
	auto dictionary_entries     = detail::required<const uncompressed_type*>(arguments, "dictionary_entries");
	auto length                 = detail::required<size_type>(arguments, "length");
	auto num_dictionary_entries = detail::required<dictionary_size_type>(arguments, "num_dictionary_entries");
	enqueue_launch(stream, launch_config, decompressed, compressed_input, dictionary_entries, length, num_dictionary_entries, caching, serialization_factor);
}

template <unsigned I, unsigned U, unsigned DI>
const device_function_t kernel_t<I, U, DI>::get_device_function() const { return giddy_b200_dictionary_kernel(I, U, DI); }

} // namespace dictionary

// =================================================================================================
// Discard zero bytes -- reference src/kernel_wrappers/decompression/discard_zero_bytes_fixed.cu:21-127,
//                                 src/kernel_wrappers/decompression/discard_zero_bytes_variable.cu:13-121
// =================================================================================================
namespace discard_zero_bytes {

namespace fixed_width {

enum : serialization_factor_t { DefaultSerializationFactor = 32 };

template <unsigned IndexSize, unsigned UncompressedSize, unsigned CompressedSize,
	util::endianness_t UncompressedEndianness = util::endianness_t::little_endian,
	serialization_factor_t SerializationFactor = DefaultSerializationFactor>
class launch_config_resolution_params_t final : public kernels::launch_config_resolution_params_t {
public:
	using parent = kernels::launch_config_resolution_params_t;
	launch_config_resolution_params_t(device::properties_t device_properties_, ::std::size_t length_,
		optional<device_function::attributes_t> known_attributes = nullopt) :
		parent(device_properties_, detail::attributes_or_query(known_attributes, giddy_b200_discard_zero_bytes_fixed_kernel(IndexSize, UncompressedSize, CompressedSize)))
	{
		static_assert(UncompressedEndianness == util::endianness_t::little_endian, "only little-endian output is supported");
		grid_construction_resolution   = thread;
		length                         = length_;
		serialization_option           = fixed_factor;
		// The template's SerializationFactor is part of the wrapper's name (and factory key) only; the
		// kernel's own per-thread element count is what sizes the grid.
		default_serialization_factor   = (16 / UncompressedSize) * GIDDY_ELEMENTWISE_UNROLL;
		quanta.threads_in_block        = warp_size;
		upper_limits.threads_in_block  = GIDDY_ELEMENTWISE_BLOCK_THREADS;
	}
};

template <unsigned IndexSize, unsigned UncompressedSize, unsigned CompressedSize,
	util::endianness_t UncompressedEndianness = util::endianness_t::little_endian,
	serialization_factor_t SerializationFactor = DefaultSerializationFactor>
class kernel_t : public cuda::registered::kernel_t {
	static_assert(UncompressedSize > CompressedSize, "Uncompressed size isn't higher than Compressed size");
public:
	GIDDY_WRAPPER_BOILERPLATE;
	using uncompressed_type = uint_t<UncompressedSize>;
	using compressed_type   = uint_t<CompressedSize>;

	static ::std::string key()
	{
		return "decompression::discard_zero_bytes::fixed_width::kernel_t<" + detail::u(IndexSize) + ", " + detail::u(UncompressedSize) + ", " +
			detail::u(CompressedSize) + ", (util::endianness_t)" + ::std::to_string((int) UncompressedEndianness) +
			", (unsigned short)" + ::std::to_string((unsigned) SerializationFactor) + ">";
	}

	launch_configuration_t resolve_launch_configuration(
		device::properties_t device_properties, device_function::attributes_t kernel_function_attributes, ::std::size_t length,
		launch_configuration_limits_t limits = { nullopt, nullopt, nullopt }) const
	{
		launch_config_resolution_params_t<IndexSize, UncompressedSize, CompressedSize, UncompressedEndianness, SerializationFactor> params(device_properties, length, optional<device_function::attributes_t>(kernel_function_attributes));
		return kernels::resolve_launch_configuration(params, limits);
	}

	void enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config,
		uncompressed_type* decompressed, const compressed_type* compressed_input, uint_t<IndexSize> length) const
	{
		auto config = to_c(launch_config);
		cuda::detail::throw_if_error(giddy_b200_discard_zero_bytes_fixed_launch(stream, &config, IndexSize, UncompressedSize, CompressedSize,
			decompressed, compressed_input, length));
	}
};

template <unsigned I, unsigned U, unsigned C, util::endianness_t E, serialization_factor_t SF>
launch_configuration_t kernel_t<I, U, C, E, SF>::resolve_launch_configuration(
	device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
	arguments_type extra_arguments, launch_configuration_limits_t limits) const
{
	auto length = detail::required<::std::size_t>(extra_arguments, "length");
	return resolve_launch_configuration(device_properties, kernel_function_attributes, length, limits);
}

template <unsigned I, unsigned U, unsigned C, util::endianness_t E, serialization_factor_t SF>
void kernel_t<I, U, C, E, SF>::enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config, arguments_type arguments) const
{
	auto decompressed     = detail::required<uncompressed_type*>(arguments, "decompressed");
	auto compressed_input = detail::required<const compressed_type*>(arguments, "compressed_input");
	auto length           = detail::required<uint_t<I>>(arguments, "length");
	enqueue_launch(stream, launch_config, decompressed, compressed_input, length);
}

template <unsigned I, unsigned U, unsigned C, util::endianness_t E, serialization_factor_t SF>
const device_function_t kernel_t<I, U, C, E, SF>::get_device_function() const { return giddy_b200_discard_zero_bytes_fixed_kernel(I, U, C); }

} // namespace fixed_width

namespace variable_width {

using element_size_t = unsigned;

template <unsigned IndexSize, unsigned UncompressedSize, unsigned ElementSizesContainerSize>
class launch_config_resolution_params_t final : public kernels::launch_config_resolution_params_t {
public:
	using parent = kernels::launch_config_resolution_params_t;
	launch_config_resolution_params_t(device::properties_t device_properties_, ::std::size_t length_in_elements, ::std::size_t segment_length,
		optional<device_function::attributes_t> known_attributes = nullopt) :
		parent(device_properties_, detail::attributes_or_query(known_attributes, giddy_b200_discard_zero_bytes_variable_kernel(IndexSize, UncompressedSize, ElementSizesContainerSize)))
	{
		if (segment_length == 0) { throw ::std::invalid_argument("The segment length must be a positive number"); }
		detail::configure_warp_per_unit(*this, util::div_rounding_up(length_in_elements, segment_length), GIDDY_SEGMENT_WARPS_PER_BLOCK);
		// each warp stages tiles of its segment (bytes via TMA, offsets, size fields) in shared memory
		dynamic_shared_memory_requirement.per_length_unit = GIDDY_DZBV_SHARED_BYTES_PER_WARP(UncompressedSize) / warp_size;
	}
};

template <unsigned IndexSize, unsigned UncompressedSize, unsigned ElementSizesContainerSize>
class kernel_t : public cuda::registered::kernel_t {
public:
	GIDDY_WRAPPER_BOILERPLATE;
	using uncompressed_type                  = uint_t<UncompressedSize>;
	using element_sizes_container_size_type  = uint_t<ElementSizesContainerSize>;
	using index_type                         = uint_t<IndexSize>;

	static ::std::string key()
	{
		return "decompression::discard_zero_bytes::variable_width::kernel_t<" + detail::u(IndexSize) + ", " + detail::u(UncompressedSize) + ", " +
			detail::u(ElementSizesContainerSize) + ">";
	}

	launch_configuration_t resolve_launch_configuration(
		device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
		::std::size_t length, ::std::size_t segment_length,
		launch_configuration_limits_t limits = { nullopt, nullopt, nullopt }) const
	{
		launch_config_resolution_params_t<IndexSize, UncompressedSize, ElementSizesContainerSize> params(device_properties, length, segment_length, optional<device_function::attributes_t>(kernel_function_attributes));
		return kernels::resolve_launch_configuration(params, limits);
	}

	void enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config,
		uncompressed_type* decompressed, const unsigned char* compressed_data, const element_sizes_container_size_type* packed_element_sizes,
		const index_type* position_anchors, index_type segment_length, index_type length,
		element_size_t min_represented_element_size, unsigned bits_per_element_size) const
	{
		auto config = to_c(launch_config);
		cuda::detail::throw_if_error(giddy_b200_discard_zero_bytes_variable_launch(stream, &config,
			IndexSize, UncompressedSize, ElementSizesContainerSize, decompressed, compressed_data, packed_element_sizes, position_anchors,
			segment_length, length, min_represented_element_size, bits_per_element_size));
	}
};

template <unsigned I, unsigned U, unsigned CS>
launch_configuration_t kernel_t<I, U, CS>::resolve_launch_configuration(
	device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
	arguments_type extra_arguments, launch_configuration_limits_t limits) const
{
	auto length         = detail::required<::std::size_t>(extra_arguments, "length");
	auto segment_length = detail::required<::std::size_t>(extra_arguments, "segment_length");
	return resolve_launch_configuration(device_properties, kernel_function_attributes, length, segment_length, limits);
}

template <unsigned I, unsigned U, unsigned CS>
void kernel_t<I, U, CS>::enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config, arguments_type arguments) const
{
	auto decompressed                 = detail::required<uncompressed_type*>(arguments, "decompressed");
	auto compressed_data              = detail::required<const unsigned char*>(arguments, "compressed_data");
	auto packed_element_sizes         = detail::required<const element_sizes_container_size_type*>(arguments, "packed_element_sizes");
	auto position_anchors             = detail::required<const index_type*>(arguments, "position_anchors");
	auto segment_length               = detail::required<index_type>(arguments, "segment_length");
	auto length                       = detail::required<index_type>(arguments, "length");
	auto min_represented_element_size = detail::required<element_size_t>(arguments, "min_represented_element_size");
	auto bits_per_element_size        = detail::required<unsigned>(arguments, "bits_per_element_size");
	enqueue_launch(stream, launch_config, decompressed, compressed_data, packed_element_sizes, position_anchors,
		segment_length, length, min_represented_element_size, bits_per_element_size);
}

template <unsigned I, unsigned U, unsigned CS>
const device_function_t kernel_t<I, U, CS>::get_device_function() const { return giddy_b200_discard_zero_bytes_variable_kernel(I, U, CS); }

} // namespace variable_width
} // namespace discard_zero_bytes

// =================================================================================================
// Extensions: fused composite decoders (keys under decompression::composite::; no reference counterpart)
// =================================================================================================
namespace composite {

/// Delta applied to the output of DZB-variable in one pass (BASELINE.json configs[3]).  Argument maps: the union
/// of the two schemes' (reference delta.cu:55-56,73-77 and discard_zero_bytes_variable.cu:58-61,80-89), with
/// "position_anchors" the byte-offset anchors of the DZB-variable stage and "anchor_values" those of the Delta stage.
namespace delta_of_discard_zero_bytes_variable_width {

using element_size_t = unsigned;

template <unsigned IndexSize, unsigned UncompressedSize, unsigned CompressedSize, unsigned ElementSizesContainerSize>
class launch_config_resolution_params_t final : public kernels::launch_config_resolution_params_t {
public:
	using parent = kernels::launch_config_resolution_params_t;
	launch_config_resolution_params_t(device::properties_t device_properties_, ::std::size_t length_in_elements, ::std::size_t segment_length,
		optional<device_function::attributes_t> known_attributes = nullopt) :
		parent(device_properties_, detail::attributes_or_query(known_attributes,
			giddy_b200_delta_of_discard_zero_bytes_variable_kernel(IndexSize, UncompressedSize, CompressedSize, ElementSizesContainerSize)))
	{
		if (segment_length == 0) { throw ::std::invalid_argument("The segment length must be a positive number"); }
		detail::configure_warp_per_unit(*this, util::div_rounding_up(length_in_elements, segment_length), GIDDY_SEGMENT_WARPS_PER_BLOCK);
		dynamic_shared_memory_requirement.per_length_unit = GIDDY_DZBV_SHARED_BYTES_PER_WARP(CompressedSize) / warp_size;
	}
};

template <unsigned IndexSize, unsigned UncompressedSize, unsigned CompressedSize, unsigned ElementSizesContainerSize>
class kernel_t : public cuda::registered::kernel_t {
public:
	GIDDY_WRAPPER_BOILERPLATE;
	using uncompressed_type                  = uint_t<UncompressedSize>;
	using element_sizes_container_size_type  = uint_t<ElementSizesContainerSize>;
	using index_type                         = uint_t<IndexSize>;

	static ::std::string key()
	{
		return "decompression::composite::delta_of_discard_zero_bytes_variable_width::kernel_t<" + detail::u(IndexSize) + ", " +
			detail::u(UncompressedSize) + ", " + detail::u(CompressedSize) + ", " + detail::u(ElementSizesContainerSize) + ">";
	}

	launch_configuration_t resolve_launch_configuration(
		device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
		::std::size_t length, ::std::size_t segment_length,
		launch_configuration_limits_t limits = { nullopt, nullopt, nullopt }) const
	{
		launch_config_resolution_params_t<IndexSize, UncompressedSize, CompressedSize, ElementSizesContainerSize> params(
			device_properties, length, segment_length, optional<device_function::attributes_t>(kernel_function_attributes));
		return kernels::resolve_launch_configuration(params, limits);
	}

	void enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config,
		uncompressed_type* decompressed, const unsigned char* compressed_data, const element_sizes_container_size_type* packed_element_sizes,
		const index_type* position_anchors, const uncompressed_type* anchor_values, index_type segment_length, index_type length,
		element_size_t min_represented_element_size, unsigned bits_per_element_size) const
	{
		auto config = to_c(launch_config);
		cuda::detail::throw_if_error(giddy_b200_delta_of_discard_zero_bytes_variable_launch(stream, &config,
			IndexSize, UncompressedSize, CompressedSize, ElementSizesContainerSize, decompressed, compressed_data, packed_element_sizes,
			position_anchors, anchor_values, segment_length, length, min_represented_element_size, bits_per_element_size));
	}
};

template <unsigned I, unsigned U, unsigned C, unsigned CS>
launch_configuration_t kernel_t<I, U, C, CS>::resolve_launch_configuration(
	device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
	arguments_type extra_arguments, launch_configuration_limits_t limits) const
{
	auto length         = detail::required<::std::size_t>(extra_arguments, "length");
	auto segment_length = detail::required<::std::size_t>(extra_arguments, "segment_length");
	return resolve_launch_configuration(device_properties, kernel_function_attributes, length, segment_length, limits);
}

template <unsigned I, unsigned U, unsigned C, unsigned CS>
void kernel_t<I, U, C, CS>::enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config, arguments_type arguments) const
{
	auto decompressed                 = detail::required<uncompressed_type*>(arguments, "decompressed");
	auto compressed_data              = detail::required<const unsigned char*>(arguments, "compressed_data");
	auto packed_element_sizes         = detail::required<const element_sizes_container_size_type*>(arguments, "packed_element_sizes");
	auto position_anchors             = detail::required<const index_type*>(arguments, "position_anchors");
	auto anchor_values                = detail::required<const uncompressed_type*>(arguments, "anchor_values");
	auto segment_length               = detail::required<index_type>(arguments, "segment_length");
	auto length                       = detail::required<index_type>(arguments, "length");
	auto min_represented_element_size = detail::required<element_size_t>(arguments, "min_represented_element_size");
	auto bits_per_element_size        = detail::required<unsigned>(arguments, "bits_per_element_size");
	enqueue_launch(stream, launch_config, decompressed, compressed_data, packed_element_sizes, position_anchors, anchor_values,
		segment_length, length, min_represented_element_size, bits_per_element_size);
}

template <unsigned I, unsigned U, unsigned C, unsigned CS>
const device_function_t kernel_t<I, U, C, CS>::get_device_function() const
{
	return giddy_b200_delta_of_discard_zero_bytes_variable_kernel(I, U, C, CS);
}

} // namespace delta_of_discard_zero_bytes_variable_width

namespace detail_patched {

/// The launch-configuration policy shared by the patched decoders: warp tiles of 32 x GIDDY_PATCHED_UNROLL output vectors,
/// handed grid-stride to the warps of a grid of GIDDY_PATCHED_BLOCKS_PER_MULTIPROCESSOR CTAs of GIDDY_ELEMENTWISE_BLOCK_
/// THREADS threads per multiprocessor (fewer when the column is short), so that a warp decodes several tiles and can
/// request a tile's patch range an iteration ahead.  The kernels are correct under any grid and any whole-warp block
/// up to that size.
inline void configure(launch_config_resolution_params_t& p, ::std::size_t length_in_elements, unsigned uncompressed_size)
{
	const ::std::size_t elements_per_block_tile = (::std::size_t) GIDDY_ELEMENTWISE_BLOCK_THREADS * GIDDY_PATCHED_UNROLL * (32 / uncompressed_size);
	p.grid_construction_resolution  = launch_config_resolution_params_t::block;
	p.serialization_option          = launch_config_resolution_params_t::auto_maximized;
	p.length                        = util::div_rounding_up(length_in_elements, elements_per_block_tile);
	p.quanta.threads_in_block       = warp_size;
	p.upper_limits.threads_in_block = GIDDY_ELEMENTWISE_BLOCK_THREADS;
	p.block_resolution_constraints.max_threads_per_block = GIDDY_ELEMENTWISE_BLOCK_THREADS;
	// blocks = multiprocessors x warp schedulers (4) x factor / warps per block
	p.keep_gpu_busy_factor          = GIDDY_PATCHED_BLOCKS_PER_MULTIPROCESSOR * (GIDDY_ELEMENTWISE_BLOCK_THREADS / warp_size) / 4;
}

} // namespace detail_patched

/// Dictionary decode and naive patching in one pass (BASELINE.json configs[2]): the argument maps of
/// decompression::dictionary (reference dictionary.cu:57-60,77-90) plus the patches of scatter (scatter.cu:55,70-73)
/// under the names "patch_data", "patch_indices" (ascending positions of type uint_t<IndexSize>) and "num_patches".
namespace patched_dictionary {

template <unsigned IndexSize, unsigned UncompressedSize, unsigned DictionaryIndexSize>
class launch_config_resolution_params_t final : public kernels::launch_config_resolution_params_t {
public:
	using parent = kernels::launch_config_resolution_params_t;
	launch_config_resolution_params_t(device::properties_t device_properties_, ::std::size_t num_dictionary_entries,
		::std::size_t data_length, optional<bool> cache_dictionary_in_shared_memory_override = nullopt,
		optional<shared_memory_size_t> dynamic_shared_mem_limit = nullopt,
		optional<device_function::attributes_t> known_attributes = nullopt) :
		parent(device_properties_, detail::attributes_or_query(known_attributes,
			giddy_b200_patched_dictionary_kernel(IndexSize, UncompressedSize, DictionaryIndexSize)), dynamic_shared_mem_limit)
	{
		const ::std::size_t dictionary_bytes = num_dictionary_entries * UncompressedSize;
		const bool worthwhile = dictionary_bytes <= dictionary::max_default_cached_dictionary_bytes &&
			dictionary_bytes <= available_dynamic_shared_memory_per_block && data_length * UncompressedSize >= 64 * dictionary_bytes;
		detail_patched::configure(*this, data_length, UncompressedSize);
		if (cache_dictionary_in_shared_memory_override.value_or(worthwhile)) {
			dynamic_shared_memory_requirement.per_block = (shared_memory_size_t) dictionary_bytes;
		}
	}
};

template <unsigned IndexSize, unsigned UncompressedSize, unsigned DictionaryIndexSize>
class kernel_t : public cuda::registered::kernel_t {
public:
	GIDDY_WRAPPER_BOILERPLATE;
	using dictionary_index_type = uint_t<DictionaryIndexSize>;
	using dictionary_size_type  = size_type_by_index_size<DictionaryIndexSize>;
	using uncompressed_type     = uint_t<UncompressedSize>;
	using index_type            = uint_t<IndexSize>;
	using size_type             = size_type_by_index_size<IndexSize>;

	static ::std::string key()
	{
		return "decompression::composite::patched_dictionary::kernel_t<" + detail::u(IndexSize) + ", " + detail::u(UncompressedSize) + ", " +
			detail::u(DictionaryIndexSize) + ">";
	}

	launch_configuration_t resolve_launch_configuration(
		device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
		::std::size_t num_dictionary_entries, ::std::size_t data_length, optional<bool> cache_input_data_in_shared_mem = nullopt,
		launch_configuration_limits_t limits = { nullopt, nullopt, nullopt }) const
	{
		launch_config_resolution_params_t<IndexSize, UncompressedSize, DictionaryIndexSize> params(device_properties, num_dictionary_entries,
			data_length, cache_input_data_in_shared_mem, limits.dynamic_shared_memory, optional<device_function::attributes_t>(kernel_function_attributes));
		return kernels::resolve_launch_configuration(params, limits);
	}

	void enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config,
		uncompressed_type* decompressed, const dictionary_index_type* compressed_input, const uncompressed_type* dictionary_entries,
		size_type length, dictionary_size_type num_dictionary_entries, bool cache_dictionary_in_shared_mem,
		const uncompressed_type* patch_data, const index_type* patch_indices, size_type num_patches) const
	{
		auto config = to_c(launch_config);
		cuda::detail::throw_if_error(giddy_b200_patched_dictionary_launch(stream, &config, IndexSize, UncompressedSize, DictionaryIndexSize,
			decompressed, compressed_input, dictionary_entries, length, num_dictionary_entries, cache_dictionary_in_shared_mem,
			patch_data, patch_indices, num_patches));
	}
};

template <unsigned I, unsigned U, unsigned DI>
launch_configuration_t kernel_t<I, U, DI>::resolve_launch_configuration(
	device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
	arguments_type extra_arguments, launch_configuration_limits_t limits) const
{
	auto num_dictionary_entries         = detail::required<::std::size_t>(extra_arguments, "num_dictionary_entries");
	auto data_length                    = detail::required<::std::size_t>(extra_arguments, "data_length");
	auto cache_input_data_in_shared_mem = detail::optional_argument<bool>(extra_arguments, "cache_input_data_in_shared_mem");
	return resolve_launch_configuration(device_properties, kernel_function_attributes, num_dictionary_entries, data_length,
		cache_input_data_in_shared_mem, limits);
}

template <unsigned I, unsigned U, unsigned DI>
void kernel_t<I, U, DI>::enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config, arguments_type arguments) const
{
	auto caching = detail::optional_argument<bool>(arguments, "cache_dictionary_in_shared_mem").value_or(launch_config.dynamic_shared_memory_size > 0);
	auto decompressed           = detail::required<uncompressed_type*>(arguments, "decompressed");
	auto compressed_input       = detail::required<const dictionary_index_type*>(arguments, "compressed_input");
	auto dictionary_entries     = detail::required<const uncompressed_type*>(arguments, "dictionary_entries");
	auto length                 = detail::required<size_type>(arguments, "length");
	auto num_dictionary_entries = detail::required<dictionary_size_type>(arguments, "num_dictionary_entries");
	auto patch_data             = detail::required<const uncompressed_type*>(arguments, "patch_data");
	auto patch_indices          = detail::required<const index_type*>(arguments, "patch_indices");
	auto num_patches            = detail::required<size_type>(arguments, "num_patches");
	enqueue_launch(stream, launch_config, decompressed, compressed_input, dictionary_entries, length, num_dictionary_entries, caching,
		patch_data, patch_indices, num_patches);
}

template <unsigned I, unsigned U, unsigned DI>
const device_function_t kernel_t<I, U, DI>::get_device_function() const { return giddy_b200_patched_dictionary_kernel(I, U, DI); }

} // namespace patched_dictionary

/// Frame-of-Reference decode and compressed-indices patching in one pass (BASELINE.json configs[4], column A): the
/// argument maps of decompression::frame_of_reference (reference frame_of_reference.cu:64-65,85-89) plus those of
/// scatter::compressed_indices (compressed_indices_scatter.cu:54-55,76-88) minus "target", with
/// OutputIndexSize = IndexSize and RunLengthSize = InputIndexSize.  Patch positions must be ascending.
namespace patched_frame_of_reference {

template <unsigned IndexSize, typename Uncompressed, typename Compressed, typename UnaryModelFunction, unsigned InputIndexSize>
inline const void* kernel_address()
{
	return giddy_b200_patched_frame_of_reference_kernel(IndexSize, sizeof(Uncompressed), sizeof(Compressed),
		(giddy_model_t) UnaryModelFunction::kind, UnaryModelFunction::index_size, InputIndexSize);
}

template <unsigned IndexSize, typename Uncompressed, typename Compressed, typename UnaryModelFunction, unsigned InputIndexSize = IndexSize>
class launch_config_resolution_params_t final : public kernels::launch_config_resolution_params_t {
public:
	using parent = kernels::launch_config_resolution_params_t;
	launch_config_resolution_params_t(device::properties_t device_properties_, ::std::size_t data_length, ::std::size_t segment_length,
		optional<device_function::attributes_t> known_attributes = nullopt) :
		parent(device_properties_, detail::attributes_or_query(known_attributes,
			kernel_address<IndexSize, Uncompressed, Compressed, UnaryModelFunction, InputIndexSize>()))
	{
		// the same preconditions as the reference, frame_of_reference.cuh:99-104
		if (segment_length == 0) { throw ::std::invalid_argument("The segment length must be a positive number"); }
		if (segment_length % sizeof(native_word_t) != 0) {
			throw ::std::invalid_argument("The segment length must be aligned with the GPU's native word size");
		}
		detail_patched::configure(*this, data_length, sizeof(Uncompressed));
	}
};

template <unsigned IndexSize, typename Uncompressed, typename Compressed, typename UnaryModelFunction, unsigned InputIndexSize = IndexSize>
class kernel_t : public cuda::registered::kernel_t {
public:
	GIDDY_WRAPPER_BOILERPLATE;
	using index_type              = uint_t<IndexSize>;
	using input_index_type        = uint_t<InputIndexSize>;
	using input_size_type         = size_type_by_index_size<InputIndexSize>;
	using model_coefficients_type = typename UnaryModelFunction::coefficients_type;

	static ::std::string key()
	{
		return "decompression::composite::patched_frame_of_reference::kernel_t<" + detail::u(IndexSize) + ", " +
			detail::type_name_of<Uncompressed>::get() + ", " + detail::type_name_of<Compressed>::get() + ", " +
			detail::type_name_of<UnaryModelFunction>::get() + ", " + detail::u(InputIndexSize) + ">";
	}

	launch_configuration_t resolve_launch_configuration(
		device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
		::std::size_t length, ::std::size_t segment_length,
		launch_configuration_limits_t limits = { nullopt, nullopt, nullopt }) const
	{
		launch_config_resolution_params_t<IndexSize, Uncompressed, Compressed, UnaryModelFunction, InputIndexSize> params(
			device_properties, length, segment_length, optional<device_function::attributes_t>(kernel_function_attributes));
		return kernels::resolve_launch_configuration(params, limits);
	}

	void enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config,
		Uncompressed* decompressed, const Compressed* compressed_input, const model_coefficients_type* segment_model_coefficients,
		index_type length, index_type segment_length,
		const Uncompressed* data_to_scatter, const input_index_type* scatter_position_run_lengths,
		const unsigned char* scatter_position_run_individual_offset_sizes, const index_type* scatter_position_run_baseline_values,
		const input_index_type* scatter_position_run_offsets_start_pos, const unsigned char* scatter_position_offset_bytes,
		const input_index_type* scatter_position_anchors, input_index_type anchoring_period, input_index_type num_scatter_position_runs,
		input_size_type input_data_length) const
	{
		auto config = to_c(launch_config);
		cuda::detail::throw_if_error(giddy_b200_patched_frame_of_reference_launch(stream, &config,
			IndexSize, sizeof(Uncompressed), sizeof(Compressed), (giddy_model_t) UnaryModelFunction::kind, UnaryModelFunction::index_size, InputIndexSize,
			decompressed, compressed_input, segment_model_coefficients, length, segment_length,
			data_to_scatter, scatter_position_run_lengths, scatter_position_run_individual_offset_sizes, scatter_position_run_baseline_values,
			scatter_position_run_offsets_start_pos, scatter_position_offset_bytes, scatter_position_anchors,
			anchoring_period, num_scatter_position_runs, input_data_length));
	}
};

template <unsigned I, typename U, typename C, typename M, unsigned II>
launch_configuration_t kernel_t<I, U, C, M, II>::resolve_launch_configuration(
	device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
	arguments_type extra_arguments, launch_configuration_limits_t limits) const
{
	auto length         = detail::required<::std::size_t>(extra_arguments, "length");
	auto segment_length = detail::required<::std::size_t>(extra_arguments, "segment_length");
	return resolve_launch_configuration(device_properties, kernel_function_attributes, length, segment_length, limits);
}

template <unsigned I, typename U, typename C, typename M, unsigned II>
void kernel_t<I, U, C, M, II>::enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config, arguments_type arguments) const
{
	auto decompressed               = detail::required<U*>(arguments, "decompressed");
	auto compressed_input           = detail::required<const C*>(arguments, "compressed_input");
	auto segment_model_coefficients = detail::required<const model_coefficients_type*>(arguments, "segment_model_coefficients");
	auto length                     = detail::required<index_type>(arguments, "length");
	auto segment_length             = detail::required<index_type>(arguments, "segment_length");
	auto data_to_scatter                              = detail::required<const U*>(arguments, "data_to_scatter");
	auto scatter_position_run_lengths                 = detail::required<const input_index_type*>(arguments, "scatter_position_run_lengths");
	auto scatter_position_run_individual_offset_sizes = detail::required<const unsigned char*>(arguments, "scatter_position_run_individual_offset_sizes");
	auto scatter_position_run_baseline_values         = detail::required<const index_type*>(arguments, "scatter_position_run_baseline_values");
	auto scatter_position_run_offsets_start_pos       = detail::required<const input_index_type*>(arguments, "scatter_position_run_offsets_start_pos");
	auto scatter_position_offset_bytes                = detail::required<const unsigned char*>(arguments, "scatter_position_offset_bytes");
	auto scatter_position_anchors                     = detail::required<const input_index_type*>(arguments, "scatter_position_anchors");
	auto anchoring_period                             = detail::required<input_index_type>(arguments, "anchoring_period");
	auto num_scatter_position_runs                    = detail::required<input_index_type>(arguments, "num_scatter_position_runs");
	auto input_data_length                            = detail::required<input_size_type>(arguments, "input_data_length");
	enqueue_launch(stream, launch_config, decompressed, compressed_input, segment_model_coefficients, length, segment_length,
		data_to_scatter, scatter_position_run_lengths, scatter_position_run_individual_offset_sizes, scatter_position_run_baseline_values,
		scatter_position_run_offsets_start_pos, scatter_position_offset_bytes, scatter_position_anchors,
		anchoring_period, num_scatter_position_runs, input_data_length);
}

template <unsigned I, typename U, typename C, typename M, unsigned II>
const device_function_t kernel_t<I, U, C, M, II>::get_device_function() const { return kernel_address<I, U, C, M, II>(); }

} // namespace patched_frame_of_reference
} // namespace composite

// =================================================================================================
// Incidence bitmaps -- reference src/kernel_wrappers/decompression/incidence_bitmaps.cu:13-102
// =================================================================================================
namespace incidence_bitmaps {

using bit_container_t = unsigned;

template <unsigned IndexSize, unsigned UncompressedSize, unsigned BitmapAlignmentInInts, bool UseDictionary = true>
class launch_config_resolution_params_t final : public kernels::launch_config_resolution_params_t {
public:
	using parent = kernels::launch_config_resolution_params_t;
	launch_config_resolution_params_t(device::properties_t device_properties_, ::std::size_t uncompressed_length,
		optional<shared_memory_size_t> = nullopt,
		optional<device_function::attributes_t> known_attributes = nullopt) :
		parent(device_properties_, detail::attributes_or_query(known_attributes, giddy_b200_incidence_bitmaps_kernel(IndexSize, UncompressedSize, BitmapAlignmentInInts, UseDictionary)))
	{
		const ::std::size_t bitmap_length_in_containers = util::div_rounding_up<::std::size_t>(uncompressed_length, 32);
		detail::configure_warp_per_unit(*this,
			util::div_rounding_up<::std::size_t>(bitmap_length_in_containers, GIDDY_BITMAPS_CONTAINERS_PER_TILE), GIDDY_SEGMENT_WARPS_PER_BLOCK);
	}
};

template <unsigned IndexSize, unsigned UncompressedSize, unsigned BitmapAlignmentInInts, bool UseDictionary = true>
class kernel_t : public cuda::registered::kernel_t {
public:
	GIDDY_WRAPPER_BOILERPLATE;
	using uncompressed_type = uint_t<UncompressedSize>;
	using size_type         = size_type_by_index_size<IndexSize>;
	using num_bitmaps_type  = size_type_by_index_size<1>;

	static ::std::string key()
	{
		return "decompression::incidence_bitmaps::kernel_t<" + detail::u(IndexSize) + ", " + detail::u(UncompressedSize) + ", " +
			detail::u(BitmapAlignmentInInts) + ", " + (UseDictionary ? "true" : "false") + ">";
	}

	launch_configuration_t resolve_launch_configuration(
		device::properties_t device_properties, device_function::attributes_t kernel_function_attributes, ::std::size_t length,
		launch_configuration_limits_t limits = { nullopt, nullopt, nullopt }) const
	{
		launch_config_resolution_params_t<IndexSize, UncompressedSize, BitmapAlignmentInInts, UseDictionary> params(device_properties, length, nullopt, optional<device_function::attributes_t>(kernel_function_attributes));
		return kernels::resolve_launch_configuration(params, limits);
	}

	void enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config,
		uncompressed_type* decompressed, const bit_container_t* incidence_bitmaps, const uncompressed_type* dictionary_entries,
		size_type bitmap_length, num_bitmaps_type num_bitmaps) const
	{
		auto config = to_c(launch_config);
		cuda::detail::throw_if_error(giddy_b200_incidence_bitmaps_launch(stream, &config,
			IndexSize, UncompressedSize, BitmapAlignmentInInts, UseDictionary, decompressed, incidence_bitmaps, dictionary_entries, bitmap_length, num_bitmaps));
	}
};

template <unsigned I, unsigned U, unsigned A, bool D>
launch_configuration_t kernel_t<I, U, A, D>::resolve_launch_configuration(
	device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
	arguments_type extra_arguments, launch_configuration_limits_t limits) const
{
	auto length = detail::required<::std::size_t>(extra_arguments, "length");
	return resolve_launch_configuration(device_properties, kernel_function_attributes, length, limits);
}

template <unsigned I, unsigned U, unsigned A, bool D>
void kernel_t<I, U, A, D>::enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config, arguments_type arguments) const
{
	auto decompressed       = detail::required<uncompressed_type*>(arguments, "decompressed");
	auto incidence_bitmaps  = detail::required<const bit_container_t*>(arguments, "incidence_bitmaps");
	auto dictionary_entries = detail::required<const uncompressed_type*>(arguments, "dictionary_entries");
	auto bitmap_length      = detail::required<size_type>(arguments, "bitmap_length");
	auto num_bitmaps        = detail::required<num_bitmaps_type>(arguments, "num_bitmaps");
	enqueue_launch(stream, launch_config, decompressed, incidence_bitmaps, dictionary_entries, bitmap_length, num_bitmaps);
}

template <unsigned I, unsigned U, unsigned A, bool D>
const device_function_t kernel_t<I, U, A, D>::get_device_function() const { return giddy_b200_incidence_bitmaps_kernel(I, U, A, D); }

} // namespace incidence_bitmaps

} // namespace decompression

// =================================================================================================
// Scatter -- reference src/kernel_wrappers/data_layout/scatter.cu:13-126
// =================================================================================================
namespace scatter {

template <unsigned OutputIndexSize, unsigned ElementSize, unsigned InputIndexSize, serialization_factor_t SerializationFactor = 16>
class launch_config_resolution_params_t final : public kernels::launch_config_resolution_params_t {
public:
	using parent = kernels::launch_config_resolution_params_t;
	launch_config_resolution_params_t(device::properties_t device_properties_, ::std::size_t input_data_length_,
		optional<device_function::attributes_t> known_attributes = nullopt) :
		parent(device_properties_, detail::attributes_or_query(known_attributes, giddy_b200_scatter_kernel(OutputIndexSize, ElementSize, InputIndexSize)))
	{
		grid_construction_resolution   = thread;
		length                         = input_data_length_;
		serialization_option           = fixed_factor;
		default_serialization_factor   = GIDDY_SCATTER_UNROLL; // the template's factor only names the wrapper
		quanta.threads_in_block        = warp_size;
		upper_limits.threads_in_block  = GIDDY_SCATTER_BLOCK_THREADS;
	}
};

template <unsigned OutputIndexSize, unsigned ElementSize, unsigned InputIndexSize, serialization_factor_t SerializationFactor = 16>
class kernel_t : public cuda::registered::kernel_t {
public:
	GIDDY_WRAPPER_BOILERPLATE;
	using element_type      = uint_t<ElementSize>;
	using input_size_type   = size_type_by_index_size<InputIndexSize>;
	using output_index_type = uint_t<OutputIndexSize>;

	static ::std::string key()
	{
		return "scatter::kernel_t<" + detail::u(OutputIndexSize) + ", " + detail::u(ElementSize) + ", " + detail::u(InputIndexSize) +
			", (unsigned short)" + ::std::to_string((unsigned) SerializationFactor) + ">";
	}

	launch_configuration_t resolve_launch_configuration(
		device::properties_t device_properties, device_function::attributes_t kernel_function_attributes, ::std::size_t data_length,
		launch_configuration_limits_t limits = { nullopt, nullopt, nullopt }) const
	{
		launch_config_resolution_params_t<OutputIndexSize, ElementSize, InputIndexSize, SerializationFactor> params(device_properties, data_length, optional<device_function::attributes_t>(kernel_function_attributes));
		return kernels::resolve_launch_configuration(params, limits);
	}

	void enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config,
		element_type* target, const element_type* data, const output_index_type* indices, input_size_type data_length) const
	{
		auto config = to_c(launch_config);
		cuda::detail::throw_if_error(giddy_b200_scatter_launch(stream, &config, OutputIndexSize, ElementSize, InputIndexSize,
			target, data, indices, data_length));
	}
};

template <unsigned OI, unsigned E, unsigned II, serialization_factor_t SF>
launch_configuration_t kernel_t<OI, E, II, SF>::resolve_launch_configuration(
	device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
	arguments_type extra_arguments, launch_configuration_limits_t limits) const
{
	auto data_length = detail::required<::std::size_t>(extra_arguments, "data_length");
	return resolve_launch_configuration(device_properties, kernel_function_attributes, data_length, limits);
}

template <unsigned OI, unsigned E, unsigned II, serialization_factor_t SF>
void kernel_t<OI, E, II, SF>::enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config, arguments_type arguments) const
{
	auto target      = detail::required<element_type*>(arguments, "target");
	auto data        = detail::required<const element_type*>(arguments, "data");
	auto indices     = detail::required<const output_index_type*>(arguments, "indices");
	auto data_length = detail::required<input_size_type>(arguments, "data_length");
	enqueue_launch(stream, launch_config, target, data, indices, data_length);
}

template <unsigned OI, unsigned E, unsigned II, serialization_factor_t SF>
const device_function_t kernel_t<OI, E, II, SF>::get_device_function() const { return giddy_b200_scatter_kernel(OI, E, II); }

// =================================================================================================
// Compressed-indices scatter -- reference src/kernel_wrappers/data_layout/compressed_indices_scatter.cu:13-153
// =================================================================================================
namespace compressed_indices {

template <unsigned OutputIndexSize, unsigned ElementSize, unsigned InputIndexSize, unsigned RunLengthSize = InputIndexSize>
class launch_config_resolution_params_t final : public kernels::launch_config_resolution_params_t {
public:
	using parent = kernels::launch_config_resolution_params_t;
	launch_config_resolution_params_t(device::properties_t device_properties_, ::std::size_t input_data_length, ::std::size_t anchoring_period,
		optional<device_function::attributes_t> known_attributes = nullopt) :
		parent(device_properties_, detail::attributes_or_query(known_attributes, giddy_b200_compressed_indices_scatter_kernel(OutputIndexSize, ElementSize, InputIndexSize, RunLengthSize)))
	{
		// as in the reference, compressed_indices_scatter.cuh:288-290
		if (input_data_length == 0) { throw ::std::invalid_argument("Zero-length scatters not currently supported"); }
		if (anchoring_period == 0) { throw ::std::invalid_argument("The anchoring period must be a positive number"); }
		detail::configure_warp_per_unit(*this, util::div_rounding_up(input_data_length, anchoring_period), GIDDY_CI_SCATTER_WARPS_PER_BLOCK);
	}
};

template <unsigned OutputIndexSize, unsigned ElementSize, unsigned InputIndexSize, unsigned RunLengthSize = InputIndexSize>
class kernel_t : public cuda::registered::kernel_t {
public:
	GIDDY_WRAPPER_BOILERPLATE;
	using element_type      = uint_t<ElementSize>;
	using input_index_type  = uint_t<InputIndexSize>;
	using input_size_type   = size_type_by_index_size<InputIndexSize>;
	using output_index_type = uint_t<OutputIndexSize>;
	using run_length_type   = uint_t<RunLengthSize>;

	static ::std::string key()
	{
		return "scatter::compressed_indices::kernel_t<" + detail::u(OutputIndexSize) + ", " + detail::u(ElementSize) + ", " +
			detail::u(InputIndexSize) + ", " + detail::u(RunLengthSize) + ">";
	}

	launch_configuration_t resolve_launch_configuration(
		device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
		::std::size_t input_data_length, ::std::size_t anchoring_period,
		launch_configuration_limits_t limits = { nullopt, nullopt, nullopt }) const
	{
		launch_config_resolution_params_t<OutputIndexSize, ElementSize, InputIndexSize, RunLengthSize> params(device_properties, input_data_length, anchoring_period, optional<device_function::attributes_t>(kernel_function_attributes));
		return kernels::resolve_launch_configuration(params, limits);
	}

	void enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config,
		element_type* target, const element_type* data_to_scatter, const run_length_type* scatter_position_run_lengths,
		const unsigned char* scatter_position_run_individual_offset_sizes, const output_index_type* scatter_position_run_baseline_values,
		const input_index_type* scatter_position_run_offsets_start_pos, const unsigned char* scatter_position_offset_bytes,
		const input_index_type* scatter_position_anchors, input_index_type anchoring_period, input_index_type num_scatter_position_runs,
		input_size_type input_data_length) const
	{
		auto config = to_c(launch_config);
		cuda::detail::throw_if_error(giddy_b200_compressed_indices_scatter_launch(stream, &config,
			OutputIndexSize, ElementSize, InputIndexSize, RunLengthSize, target, data_to_scatter, scatter_position_run_lengths,
			scatter_position_run_individual_offset_sizes, scatter_position_run_baseline_values, scatter_position_run_offsets_start_pos,
			scatter_position_offset_bytes, scatter_position_anchors, anchoring_period, num_scatter_position_runs, input_data_length));
	}
};

template <unsigned OI, unsigned E, unsigned II, unsigned RL>
launch_configuration_t kernel_t<OI, E, II, RL>::resolve_launch_configuration(
	device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
	arguments_type extra_arguments, launch_configuration_limits_t limits) const
{
	auto input_data_length = detail::required<::std::size_t>(extra_arguments, "input_data_length");
	auto anchoring_period  = detail::required<::std::size_t>(extra_arguments, "anchoring_period");
	return resolve_launch_configuration(device_properties, kernel_function_attributes, input_data_length, anchoring_period, limits);
}

template <unsigned OI, unsigned E, unsigned II, unsigned RL>
void kernel_t<OI, E, II, RL>::enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config, arguments_type arguments) const
{
	auto target                                       = detail::required<element_type*>(arguments, "target");
	auto data_to_scatter                              = detail::required<const element_type*>(arguments, "data_to_scatter");
	auto scatter_position_run_lengths                 = detail::required<const run_length_type*>(arguments, "scatter_position_run_lengths");
	auto scatter_position_run_individual_offset_sizes = detail::required<const unsigned char*>(arguments, "scatter_position_run_individual_offset_sizes");
	auto scatter_position_run_baseline_values         = detail::required<const output_index_type*>(arguments, "scatter_position_run_baseline_values");
	auto scatter_position_run_offsets_start_pos       = detail::required<const input_index_type*>(arguments, "scatter_position_run_offsets_start_pos");
	auto scatter_position_offset_bytes                = detail::required<const unsigned char*>(arguments, "scatter_position_offset_bytes");
	auto scatter_position_anchors                     = detail::required<const input_index_type*>(arguments, "scatter_position_anchors");
	auto anchoring_period                             = detail::required<input_index_type>(arguments, "anchoring_period");
	auto num_scatter_position_runs                    = detail::required<input_index_type>(arguments, "num_scatter_position_runs");
	auto input_data_length                            = detail::required<input_size_type>(arguments, "input_data_length");
	if (launch_config.grid_dimensions.empty()) { return; } // no patches to apply (reference compressed_indices_scatter.cu:69-73)
	enqueue_launch(stream, launch_config, target, data_to_scatter, scatter_position_run_lengths, scatter_position_run_individual_offset_sizes,
		scatter_position_run_baseline_values, scatter_position_run_offsets_start_pos, scatter_position_offset_bytes, scatter_position_anchors,
		anchoring_period, num_scatter_position_runs, input_data_length);
}

template <unsigned OI, unsigned E, unsigned II, unsigned RL>
const device_function_t kernel_t<OI, E, II, RL>::get_device_function() const
{
	return giddy_b200_compressed_indices_scatter_kernel(OI, E, II, RL);
}

} // namespace compressed_indices
} // namespace scatter

#undef GIDDY_WRAPPER_BOILERPLATE

} // namespace kernels
} // namespace cuda

#endif // GIDDY_KERNEL_WRAPPERS_HPP_
