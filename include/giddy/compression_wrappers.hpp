// giddy/compression_wrappers.hpp -- kernel wrappers of the compressors (include/giddy_b200.h, group 8) and of the
// utility kernels which the reference has but does not wrap (gather, per-interval models, prefix sums).
//
// The reference ships decompressors only ("compression ... is not included", README.md:128-130, its open Issue #3), so
// there are no wrapper classes to mirror here, only the reference's wrapper *interface*: a factory key, a launch
// configuration resolved from named size_t arguments, a launch with named, exactly typed arguments
// (src/cuda/kernel_wrapper.cuh:47-123).  Keys are the matching decompressor's with "decompression::" replaced by
// "compression::"; argument names are the decompressor's where the array is the same one ("run_data", "position_anchors",
// "segment_model_coefficients" ...), plus "uncompressed" for the column going in and the outputs a compressor adds.
//
// Every compressor takes, optionally, "status": unsigned* -- a device word into which GIDDY_COMPRESSION_* bits are OR-ed
// when the column does not fit the format (zero it beforehand; it is never cleared).
#pragma once
#ifndef GIDDY_COMPRESSION_WRAPPERS_HPP_
#define GIDDY_COMPRESSION_WRAPPERS_HPP_

#include "kernel_wrappers.hpp"

namespace cuda {
namespace kernels {

namespace detail {

/// launch_config_resolution_params_t with a public constructor, for wrappers described by a traits class
class described_params_t final : public kernels::launch_config_resolution_params_t {
public:
	described_params_t(device::properties_t device_properties_, device_function::attributes_t attributes) :
		kernels::launch_config_resolution_params_t(device_properties_, attributes) { }
};

/// A wrapper whose three virtual functions are supplied by a traits class:
///   static std::string key();   static const void* kernel();
///   static void configure(described_params_t&, const arguments_type&);           // from the named size_t arguments
///   static void launch(stream::id_t, const giddy_launch_config_t&, const arguments_type&);
template <typename Traits>
class described_kernel_t : public cuda::registered::kernel_t {
public:
	using parent          = cuda::registered::kernel_t;
	using arguments_type  = parent::arguments_type;
	using parent::resolve_launch_configuration;
	static ::std::string key() { return Traits::key(); }
	static void registerInSubclassFactory() { parent::template registerInSubclassFactory<described_kernel_t>(key()); }

	launch_configuration_t resolve_launch_configuration(device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
		arguments_type extra_arguments, launch_configuration_limits_t limits = { nullopt, nullopt, nullopt }) const override final
	{
		described_params_t params(device_properties, kernel_function_attributes);
		Traits::configure(params, extra_arguments);
		return kernels::resolve_launch_configuration(params, limits);
	}
	void enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config, arguments_type arguments) const override final
	{
		Traits::launch(stream, to_c(launch_config), arguments);
	}
	const device_function_t get_device_function() const override final { return Traits::kernel(); }
};

/// The streaming kernels' configuration: thread resolution, `elements_per_thread` elements per thread
inline void configure_streaming(launch_config_resolution_params_t& p, ::std::size_t length, unsigned elements_per_thread)
{
	p.grid_construction_resolution   = launch_config_resolution_params_t::thread;
	p.length                         = length;
	p.serialization_option           = launch_config_resolution_params_t::fixed_factor;
	p.default_serialization_factor   = (serialization_factor_t) elements_per_thread;
	p.quanta.threads_in_block        = warp_size;
	p.upper_limits.threads_in_block  = GIDDY_ELEMENTWISE_BLOCK_THREADS;
}

/// The tiled kernels' configuration: a block of exactly `threads` threads per tile, grid-stride over the tiles
inline void configure_tiled(launch_config_resolution_params_t& p, ::std::size_t num_tiles, unsigned threads)
{
	p.grid_construction_resolution   = launch_config_resolution_params_t::block;
	p.serialization_option           = launch_config_resolution_params_t::auto_maximized;
	p.length                         = num_tiles;
	p.quanta.threads_in_block        = warp_size;
	p.block_resolution_constraints.fixed_threads_per_block = threads;
	p.keep_gpu_busy_factor           = 8 * (threads / warp_size) / 4; // 8 blocks per multiprocessor
}

inline unsigned* status_argument(const kernel_t::arguments_type& arguments)
{
	auto it = arguments.find("status");
	return it == arguments.end() ? nullptr : ::cuda::any_cast<unsigned*>(it->second);
}

template <typename T>
inline T argument_or(const kernel_t::arguments_type& arguments, const char* name, T fallback)
{
	auto it = arguments.find(name);
	return it == arguments.end() ? fallback : ::cuda::any_cast<T>(it->second);
}

} // namespace detail

namespace compression {

// ---- Delta: format of decompression::delta (reference src/kernels/decompression/delta.cuh:66-108) -------------------
namespace delta {
template <unsigned IndexSize, unsigned UncompressedSize, unsigned CompressedSize>
struct traits {
	using U = uint_t<UncompressedSize>; using C = uint_t<CompressedSize>; using I = uint_t<IndexSize>;
	static ::std::string key() { return "compression::delta::kernel_t<" + detail::u(IndexSize) + ", " + detail::u(UncompressedSize) + ", " + detail::u(CompressedSize) + ">"; }
	static const void* kernel() { return giddy_b200_delta_compress_kernel(IndexSize, UncompressedSize, CompressedSize); }
	static void configure(detail::described_params_t& p, const kernel_t::arguments_type& a)
	{
		if (detail::required<::std::size_t>(a, "segment_length") == 0) { throw ::std::invalid_argument("The segment length must be a positive number"); }
		detail::configure_streaming(p, detail::required<::std::size_t>(a, "length"), (16 / UncompressedSize) * GIDDY_ELEMENTWISE_UNROLL);
	}
	/// "compressed_output": C*, "anchor_values": U* (one per segment), "uncompressed": const U*, "length", "segment_length": I, "base_value": U (optional, 0)
	static void launch(stream::id_t stream, const giddy_launch_config_t& config, const kernel_t::arguments_type& a)
	{
		cuda::detail::throw_if_error(giddy_b200_delta_compress_launch(stream, &config, IndexSize, UncompressedSize, CompressedSize,
			detail::required<C*>(a, "compressed_output"), detail::required<U*>(a, "anchor_values"), detail::required<const U*>(a, "uncompressed"),
			detail::required<I>(a, "length"), detail::required<I>(a, "segment_length"), (uint64_t) detail::argument_or<U>(a, "base_value", 0),
			detail::status_argument(a)));
	}
};
template <unsigned I, unsigned U, unsigned C> using kernel_t = detail::described_kernel_t<traits<I, U, C>>;
} // namespace delta

// ---- Discard zero bytes, fixed width (discard_zero_bytes_fixed.cuh:108-124) -----------------------------------------
namespace discard_zero_bytes {
namespace fixed_width {
template <unsigned IndexSize, unsigned UncompressedSize, unsigned CompressedSize>
struct traits {
	using U = uint_t<UncompressedSize>; using C = uint_t<CompressedSize>; using I = uint_t<IndexSize>;
	static ::std::string key() { return "compression::discard_zero_bytes::fixed_width::kernel_t<" + detail::u(IndexSize) + ", " + detail::u(UncompressedSize) + ", " + detail::u(CompressedSize) + ">"; }
	static const void* kernel() { return giddy_b200_discard_zero_bytes_fixed_compress_kernel(IndexSize, UncompressedSize, CompressedSize); }
	static void configure(detail::described_params_t& p, const kernel_t::arguments_type& a)
	{
		detail::configure_streaming(p, detail::required<::std::size_t>(a, "length"), (16 / UncompressedSize) * GIDDY_ELEMENTWISE_UNROLL);
	}
	/// "compressed_output": C*, "uncompressed": const U*, "length": I
	static void launch(stream::id_t stream, const giddy_launch_config_t& config, const kernel_t::arguments_type& a)
	{
		cuda::detail::throw_if_error(giddy_b200_discard_zero_bytes_fixed_compress_launch(stream, &config, IndexSize, UncompressedSize, CompressedSize,
			detail::required<C*>(a, "compressed_output"), detail::required<const U*>(a, "uncompressed"), detail::required<I>(a, "length"), detail::status_argument(a)));
	}
};
template <unsigned I, unsigned U, unsigned C> using kernel_t = detail::described_kernel_t<traits<I, U, C>>;
} // namespace fixed_width

// ---- Discard zero bytes, variable width (discard_zero_bytes_variable.cuh:335-439) ------------------------------------
namespace variable_width {
template <unsigned IndexSize, unsigned UncompressedSize, unsigned ElementSizesContainerSize>
struct traits {
	using U = uint_t<UncompressedSize>; using CS = uint_t<ElementSizesContainerSize>; using I = uint_t<IndexSize>;
	static ::std::string key() { return "compression::discard_zero_bytes::variable_width::kernel_t<" + detail::u(IndexSize) + ", " + detail::u(UncompressedSize) + ", " + detail::u(ElementSizesContainerSize) + ">"; }
	static const void* kernel() { return giddy_b200_discard_zero_bytes_variable_compress_kernel(IndexSize, UncompressedSize, ElementSizesContainerSize); }
	static void configure(detail::described_params_t& p, const kernel_t::arguments_type& a)
	{
		const auto segment_length = detail::required<::std::size_t>(a, "segment_length");
		if (segment_length == 0) { throw ::std::invalid_argument("The segment length must be a positive number"); }
		detail::configure_warp_per_unit(p, util::div_rounding_up(detail::required<::std::size_t>(a, "length"), segment_length), GIDDY_SEGMENT_WARPS_PER_BLOCK);
	}
	/// "compressed_data": unsigned char* (room for "data_capacity": size_t bytes), "packed_element_sizes": CS*, "position_anchors": I*,
	/// "num_bytes": unsigned long* (optional: the stream's length), "uncompressed": const U*, "segment_length", "length": I,
	/// "min_represented_element_size", "bits_per_element_size": unsigned
	static void launch(stream::id_t stream, const giddy_launch_config_t& config, const kernel_t::arguments_type& a)
	{
		cuda::detail::throw_if_error(giddy_b200_discard_zero_bytes_variable_compress_launch(stream, &config, IndexSize, UncompressedSize, ElementSizesContainerSize,
			detail::required<unsigned char*>(a, "compressed_data"), detail::required<CS*>(a, "packed_element_sizes"), detail::required<I*>(a, "position_anchors"),
			detail::argument_or<::std::uint64_t*>(a, "num_bytes", nullptr), detail::required<const U*>(a, "uncompressed"),
			detail::required<I>(a, "length"), detail::required<I>(a, "segment_length"),
			detail::required<unsigned>(a, "min_represented_element_size"), detail::required<unsigned>(a, "bits_per_element_size"),
			detail::required<::std::size_t>(a, "data_capacity"), detail::status_argument(a)));
	}
};
template <unsigned I, unsigned U, unsigned CS> using kernel_t = detail::described_kernel_t<traits<I, U, CS>>;
} // namespace variable_width
} // namespace discard_zero_bytes

// ---- Dictionary: index lookup against a given (ascending) dictionary (dictionary.cuh:61-79) ---------------------------
namespace dictionary {
template <unsigned IndexSize, unsigned UncompressedSize, unsigned DictionaryIndexSize>
struct traits {
	using U = uint_t<UncompressedSize>; using DI = uint_t<DictionaryIndexSize>;
	static ::std::string key() { return "compression::dictionary::kernel_t<" + detail::u(IndexSize) + ", " + detail::u(UncompressedSize) + ", " + detail::u(DictionaryIndexSize) + ">"; }
	static const void* kernel() { return giddy_b200_dictionary_compress_kernel(IndexSize, UncompressedSize, DictionaryIndexSize); }
	static void configure(detail::described_params_t& p, const kernel_t::arguments_type& a)
	{
		// every block builds its own copy of the dictionary (a hash table, or the entries for a binary search) in shared memory:
		// a few blocks per multiprocessor which stride over the column, not a block per tile
		const ::std::size_t per_tile = (::std::size_t) GIDDY_ELEMENTWISE_BLOCK_THREADS * (16 / UncompressedSize) * GIDDY_ELEMENTWISE_UNROLL;
		detail::configure_tiled(p, (detail::required<::std::size_t>(a, "data_length") + per_tile - 1) / per_tile, GIDDY_ELEMENTWISE_BLOCK_THREADS);
		// the dictionary is searched in shared memory when it fits what a block gets without asking
		const ::std::size_t bytes = detail::required<::std::size_t>(a, "num_dictionary_entries") * UncompressedSize;
		if (bytes <= 32 * 1024) { p.dynamic_shared_memory_requirement.per_block = (shared_memory_size_t) bytes; }
	}
	/// "compressed_output": DI*, "uncompressed": const U*, "dictionary_entries": const U* (ascending), "num_dictionary_entries": size_type<DI>, "length": size_type<I>
	static void launch(stream::id_t stream, const giddy_launch_config_t& config, const kernel_t::arguments_type& a)
	{
		cuda::detail::throw_if_error(giddy_b200_dictionary_compress_launch(stream, &config, IndexSize, UncompressedSize, DictionaryIndexSize,
			detail::required<DI*>(a, "compressed_output"), detail::required<const U*>(a, "uncompressed"), detail::required<const U*>(a, "dictionary_entries"),
			detail::required<size_type_by_index_size<DictionaryIndexSize>>(a, "num_dictionary_entries"),
			detail::required<size_type_by_index_size<IndexSize>>(a, "length"), detail::status_argument(a)));
	}
};
template <unsigned I, unsigned U, unsigned DI> using kernel_t = detail::described_kernel_t<traits<I, U, DI>>;
} // namespace dictionary

// ---- Frame of Reference (frame_of_reference.cuh:35-80): per-segment model fit + offsets -------------------------------
namespace frame_of_reference {
template <unsigned IndexSize, typename Uncompressed, typename Compressed, typename UnaryModelFunction>
struct traits {
	static constexpr bool linear = (giddy_model_t) UnaryModelFunction::kind == GIDDY_MODEL_LINEAR;
	using I = uint_t<IndexSize>;
	using coefficients_type = typename UnaryModelFunction::coefficients_type;
	static ::std::string key()
	{
		return "compression::frame_of_reference::kernel_t<" + detail::u(IndexSize) + ", " + detail::type_name_of<Uncompressed>::get() + ", " +
			detail::type_name_of<Compressed>::get() + ", " + detail::type_name_of<UnaryModelFunction>::get() + " >";
	}
	static const void* kernel() { return giddy_b200_frame_of_reference_compress_kernel(IndexSize, sizeof(Uncompressed), sizeof(Compressed), linear ? GIDDY_MODEL_LINEAR : GIDDY_MODEL_CONSTANT); }
	static void configure(detail::described_params_t& p, const kernel_t::arguments_type& a)
	{
		const auto segment_length = detail::required<::std::size_t>(a, "segment_length");
		if (segment_length == 0) { throw ::std::invalid_argument("The segment length must be a positive number"); }
		detail::configure_warp_per_unit(p, util::div_rounding_up(detail::required<::std::size_t>(a, "length"), segment_length), GIDDY_SEGMENT_WARPS_PER_BLOCK);
	}
	/// "compressed_output": Compressed*, "segment_model_coefficients": coefficients_type* (one per segment), "uncompressed": const Uncompressed*,
	/// "length", "segment_length": I
	static void launch(stream::id_t stream, const giddy_launch_config_t& config, const kernel_t::arguments_type& a)
	{
		cuda::detail::throw_if_error(giddy_b200_frame_of_reference_compress_launch(stream, &config, IndexSize, sizeof(Uncompressed), sizeof(Compressed),
			linear ? GIDDY_MODEL_LINEAR : GIDDY_MODEL_CONSTANT,
			detail::required<Compressed*>(a, "compressed_output"), detail::required<coefficients_type*>(a, "segment_model_coefficients"),
			detail::required<const Uncompressed*>(a, "uncompressed"), detail::required<I>(a, "length"), detail::required<I>(a, "segment_length"),
			detail::status_argument(a)));
	}
};
template <unsigned I, typename U, typename C, typename M> using kernel_t = detail::described_kernel_t<traits<I, U, C, M>>;
} // namespace frame_of_reference

// ---- Run-Length Encoding (run_length_encoding.cuh:284-374) ---------------------------------------------------------------
namespace run_length_encoding {
template <unsigned UncompressedIndexSize, unsigned UncompressedSize, unsigned RunLengthSize, unsigned RunIndexSize = UncompressedIndexSize>
struct traits {
	using U = uint_t<UncompressedSize>; using RL = uint_t<RunLengthSize>; using RI = uint_t<RunIndexSize>;
	static ::std::string key()
	{
		return "compression::run_length_encoding::kernel_t<" + detail::u(UncompressedIndexSize) + ", " + detail::u(UncompressedSize) + ", " +
			detail::u(RunLengthSize) + ", " + detail::u(RunIndexSize) + ">";
	}
	static const void* kernel() { return giddy_b200_run_length_encoding_compress_kernel(UncompressedIndexSize, UncompressedSize, RunLengthSize, RunIndexSize); }
	static void configure(detail::described_params_t& p, const kernel_t::arguments_type& a)
	{
		if (detail::required<::std::size_t>(a, "segment_length") == 0) { throw ::std::invalid_argument("The segment length must be a positive number"); }
		detail::configure_tiled(p, util::div_rounding_up(detail::required<::std::size_t>(a, "length"), (::std::size_t) 4096), 256);
	}
	/// "run_data": U*, "run_lengths": RL* (room for "run_capacity": size_t runs each), "position_anchors": RI*, "intra_run_anchor_offsets": RL*
	/// (one per segment), "num_element_runs": size_type<RI>* (optional: the number of runs), "uncompressed": const U*,
	/// "segment_length", "length": size_type<UIdx>, "max_run_length": size_t (optional; 0 = the run-length type's maximum)
	static void launch(stream::id_t stream, const giddy_launch_config_t& config, const kernel_t::arguments_type& a)
	{
		using size = size_type_by_index_size<UncompressedIndexSize>;
		cuda::detail::throw_if_error(giddy_b200_run_length_encoding_compress_launch(stream, &config, UncompressedIndexSize, UncompressedSize, RunLengthSize, RunIndexSize,
			detail::required<U*>(a, "run_data"), detail::required<RL*>(a, "run_lengths"), detail::required<RI*>(a, "position_anchors"),
			detail::required<RL*>(a, "intra_run_anchor_offsets"), detail::argument_or<size_type_by_index_size<RunIndexSize>*>(a, "num_element_runs", nullptr),
			detail::required<const U*>(a, "uncompressed"), detail::required<size>(a, "length"), detail::required<size>(a, "segment_length"),
			detail::argument_or<::std::size_t>(a, "max_run_length", 0), detail::required<::std::size_t>(a, "run_capacity"), detail::status_argument(a)));
	}
};
template <unsigned UI, unsigned U, unsigned RL, unsigned RI = UI> using kernel_t = detail::described_kernel_t<traits<UI, U, RL, RI>>;
} // namespace run_length_encoding

// ---- Run-Position Encoding (run_position_encoding.cuh:122-359) -----------------------------------------------------------
namespace run_position_encoding {
template <unsigned UncompressedIndexSize, unsigned UncompressedSize, unsigned PositionOffsetSize, bool PositionsAreRelative, unsigned RunIndexSize = UncompressedIndexSize>
struct traits {
	using U = uint_t<UncompressedSize>; using PO = uint_t<PositionOffsetSize>; using RI = uint_t<RunIndexSize>;
	static ::std::string key()
	{
		return "compression::run_position_encoding::kernel_t<" + detail::u(UncompressedIndexSize) + ", " + detail::u(UncompressedSize) + ", " +
			detail::u(PositionOffsetSize) + ", " + (PositionsAreRelative ? "true" : "false") + ", " + detail::u(RunIndexSize) + ">";
	}
	static const void* kernel() { return giddy_b200_run_position_encoding_compress_kernel(UncompressedIndexSize, UncompressedSize, PositionOffsetSize, RunIndexSize); }
	static void configure(detail::described_params_t& p, const kernel_t::arguments_type& a)
	{
		if (detail::required<::std::size_t>(a, "segment_length") == 0) { throw ::std::invalid_argument("The segment length must be a positive number"); }
		detail::configure_tiled(p, util::div_rounding_up(detail::required<::std::size_t>(a, "length"), (::std::size_t) 4096), 256);
	}
	/// "run_data": U*, "run_start_positions": PO* (room for "run_capacity": size_t runs each), "position_anchors": RI* (one per segment),
	/// "num_element_runs": size_type<RI>* (optional), "uncompressed": const U*, "segment_length", "length": size_type<UIdx>,
	/// "max_run_length": size_t (optional; 0 = runs are never cut)
	static void launch(stream::id_t stream, const giddy_launch_config_t& config, const kernel_t::arguments_type& a)
	{
		using size = size_type_by_index_size<UncompressedIndexSize>;
		cuda::detail::throw_if_error(giddy_b200_run_position_encoding_compress_launch(stream, &config, UncompressedIndexSize, UncompressedSize, PositionOffsetSize,
			PositionsAreRelative ? 1 : 0, RunIndexSize,
			detail::required<U*>(a, "run_data"), detail::required<PO*>(a, "run_start_positions"), detail::required<RI*>(a, "position_anchors"),
			detail::argument_or<size_type_by_index_size<RunIndexSize>*>(a, "num_element_runs", nullptr),
			detail::required<const U*>(a, "uncompressed"), detail::required<size>(a, "length"), detail::required<size>(a, "segment_length"),
			detail::argument_or<::std::size_t>(a, "max_run_length", 0), detail::required<::std::size_t>(a, "run_capacity"), detail::status_argument(a)));
	}
};
template <unsigned UI, unsigned U, unsigned PO, bool Rel, unsigned RI = UI> using kernel_t = detail::described_kernel_t<traits<UI, U, PO, Rel, RI>>;
} // namespace run_position_encoding

// ---- Incidence bitmaps (incidence_bitmaps.cuh:69-131) ---------------------------------------------------------------------
namespace incidence_bitmaps {
template <unsigned IndexSize, unsigned UncompressedSize, unsigned BitmapAlignmentInInts, bool UseDictionary = true>
struct traits {
	using U = uint_t<UncompressedSize>;
	static ::std::string key()
	{
		return "compression::incidence_bitmaps::kernel_t<" + detail::u(IndexSize) + ", " + detail::u(UncompressedSize) + ", " +
			detail::u(BitmapAlignmentInInts) + ", " + (UseDictionary ? "true" : "false") + ">";
	}
	static const void* kernel() { return giddy_b200_incidence_bitmaps_compress_kernel(UncompressedSize, UseDictionary ? 1 : 0); }
	static void configure(detail::described_params_t& p, const kernel_t::arguments_type& a)
	{
		detail::configure_warp_per_unit(p, util::div_rounding_up(detail::required<::std::size_t>(a, "length"), (::std::size_t) 32), GIDDY_SEGMENT_WARPS_PER_BLOCK);
		p.serialization_option = launch_config_resolution_params_t::auto_maximized;
		p.keep_gpu_busy_factor = 16;
	}
	/// "incidence_bitmaps": unsigned* (num_bitmaps x aligned(ceil(length / 32)) words), "uncompressed": const U*, "dictionary_entries": const U*,
	/// "length": size_type<I>, "num_bitmaps": unsigned
	static void launch(stream::id_t stream, const giddy_launch_config_t& config, const kernel_t::arguments_type& a)
	{
		cuda::detail::throw_if_error(giddy_b200_incidence_bitmaps_compress_launch(stream, &config, UncompressedSize, BitmapAlignmentInInts, UseDictionary ? 1 : 0,
			detail::required<unsigned*>(a, "incidence_bitmaps"), detail::required<const U*>(a, "uncompressed"), detail::required<const U*>(a, "dictionary_entries"),
			detail::required<size_type_by_index_size<IndexSize>>(a, "length"), detail::required<unsigned>(a, "num_bitmaps"), detail::status_argument(a)));
	}
};
template <unsigned I, unsigned U, unsigned A, bool D = true> using kernel_t = detail::described_kernel_t<traits<I, U, A, D>>;
} // namespace incidence_bitmaps

} // namespace compression

// =================================================================================================
// Utility kernels the reference has but neither wraps nor registers
// =================================================================================================

// ---- gather (reference src/kernels/data_layout/gather.cuh:151-223): reordered_data[i] = data[indices[i]] -----------------------
// The kernel behind dictionary::decompress; its launch-configuration parameters are the reference's own for gather (:166-223),
// which the Dictionary wrapper already uses.
namespace gather {
template <unsigned OutputIndexSize, unsigned ElementSize, unsigned InputIndexSize>
class kernel_t : public cuda::registered::kernel_t {
public:
	using parent          = cuda::registered::kernel_t;
	using arguments_type  = parent::arguments_type;
	using parent::resolve_launch_configuration;
	using element_type = uint_t<ElementSize>; using index_type = uint_t<InputIndexSize>;
	static ::std::string key() { return "gather::kernel_t<" + detail::u(OutputIndexSize) + ", " + detail::u(ElementSize) + ", " + detail::u(InputIndexSize) + ">"; }
	static void registerInSubclassFactory() { parent::template registerInSubclassFactory<kernel_t>(key()); }

	/// "data_length", "num_indices": size_t; optional "cache_input_data_in_shared_mem": bool, "serialization_factor": serialization_factor_t
	launch_configuration_t resolve_launch_configuration(device::properties_t device_properties, device_function::attributes_t kernel_function_attributes,
		arguments_type extra_arguments, launch_configuration_limits_t limits = { nullopt, nullopt, nullopt }) const override final
	{
		decompression::dictionary::launch_config_resolution_params_t<OutputIndexSize, ElementSize, InputIndexSize> params(device_properties,
			detail::required<::std::size_t>(extra_arguments, "data_length"), detail::required<::std::size_t>(extra_arguments, "num_indices"),
			detail::optional_argument<bool>(extra_arguments, "cache_input_data_in_shared_mem"), limits.dynamic_shared_memory,
			optional<device_function::attributes_t>(kernel_function_attributes));
		return kernels::resolve_launch_configuration(params, limits, detail::optional_argument<serialization_factor_t>(extra_arguments, "serialization_factor"));
	}
	/// "reordered_data": E*, "data": const E*, "indices": const II*, "data_length": size_type<II>, "num_indices": size_type<OI>;
	/// optional "cache_data_in_shared_memory": bool, "serialization_factor"
	void enqueue_launch(stream::id_t stream, const launch_configuration_t& launch_config, arguments_type a) const override final
	{
		auto config = to_c(launch_config);
		cuda::detail::throw_if_error(giddy_b200_gather_launch(stream, &config, OutputIndexSize, ElementSize, InputIndexSize,
			detail::required<element_type*>(a, "reordered_data"), detail::required<const element_type*>(a, "data"), detail::required<const index_type*>(a, "indices"),
			detail::required<size_type_by_index_size<InputIndexSize>>(a, "data_length"), detail::required<size_type_by_index_size<OutputIndexSize>>(a, "num_indices"),
			detail::optional_argument<bool>(a, "cache_data_in_shared_memory").value_or(launch_config.dynamic_shared_memory_size > 0) ? 1 : 0,
			detail::optional_argument<serialization_factor_t>(a, "serialization_factor").value_or((serialization_factor_t)((16 / ElementSize) * GIDDY_ELEMENTWISE_UNROLL))));
	}
	const device_function_t get_device_function() const override final { return giddy_b200_gather_kernel(OutputIndexSize, ElementSize, InputIndexSize); }
};
} // namespace gather

namespace decompression {
namespace model {
// ---- decompress_with_intervals (reference src/kernels/decompression/model.cuh:55-98): one model per interval ----------------
namespace intervals {
template <unsigned IndexSize, typename Uncompressed, typename UnaryModelFunction>
struct traits {
	using I = uint_t<IndexSize>;
	using coefficients_type = typename UnaryModelFunction::coefficients_type;
	static ::std::string key()
	{
		return "decompression::model::intervals::kernel_t<" + detail::u(IndexSize) + ", " + detail::type_name_of<Uncompressed>::get() + ", " +
			detail::type_name_of<UnaryModelFunction>::get() + " >";
	}
	static const void* kernel() { return giddy_b200_model_with_intervals_kernel(IndexSize, sizeof(Uncompressed), (giddy_model_t) UnaryModelFunction::kind); }
	static void configure(detail::described_params_t& p, const cuda::kernel_t::arguments_type& a)
	{
		if (detail::required<::std::size_t>(a, "modeling_period") == 0) { throw ::std::invalid_argument("The modeling period must be a positive number"); }
		detail::configure_streaming(p, detail::required<::std::size_t>(a, "length"), (16 / sizeof(Uncompressed)) * GIDDY_ELEMENTWISE_UNROLL);
	}
	/// "decompressed": Uncompressed*, "interval_model_coefficients": const coefficients_type*, "modeling_period", "length": I
	static void launch(stream::id_t stream, const giddy_launch_config_t& config, const cuda::kernel_t::arguments_type& a)
	{
		cuda::detail::throw_if_error(giddy_b200_model_with_intervals_launch(stream, &config, IndexSize, sizeof(Uncompressed), (giddy_model_t) UnaryModelFunction::kind,
			detail::required<Uncompressed*>(a, "decompressed"), detail::required<const coefficients_type*>(a, "interval_model_coefficients"),
			detail::required<I>(a, "modeling_period"), detail::required<I>(a, "length")));
	}
};
template <unsigned I, typename U, typename M> using kernel_t = detail::described_kernel_t<traits<I, U, M>>;
} // namespace intervals
} // namespace model
} // namespace decompression

// ---- inclusive prefix sum (reference src/kernels/reduction/scan.cuh:127, 412, 525 with ReductionOp = plus) ------------------
namespace reduction {
namespace scan {
template <unsigned IndexSize, unsigned ResultSize, unsigned InputSize>
struct traits {
	using R = uint_t<ResultSize>; using D = uint_t<InputSize>; using I = uint_t<IndexSize>;
	static ::std::string key() { return "reduction::scan::inclusive::plus::kernel_t<" + detail::u(IndexSize) + ", " + detail::u(ResultSize) + ", " + detail::u(InputSize) + ">"; }
	static const void* kernel() { return giddy_b200_inclusive_scan_kernel(IndexSize, ResultSize, InputSize); }
	static void configure(detail::described_params_t& p, const cuda::kernel_t::arguments_type& a)
	{
		// a single segment as long as the data: the Delta wrapper's configuration (tiles with decoupled look-back once the data is
		// long enough, one warp otherwise)
		const auto length = detail::required<::std::size_t>(a, "length");
		decompression::delta::launch_config_resolution_params_t<IndexSize, ResultSize, InputSize> as_delta(p.device_properties, length, length ? length : 1,
			nullopt, optional<device_function::attributes_t>(p.kernel_function_attributes));
		static_cast<launch_config_resolution_params_t&>(p) = as_delta;
	}
	/// "result": R*, "data": const D*, "length": I, optional "initial_value": const R* (a device pointer; zero when absent)
	static void launch(stream::id_t stream, const giddy_launch_config_t& config, const cuda::kernel_t::arguments_type& a)
	{
		cuda::detail::throw_if_error(giddy_b200_inclusive_scan_launch(stream, &config, IndexSize, ResultSize, InputSize,
			detail::required<R*>(a, "result"), detail::required<const D*>(a, "data"), detail::required<I>(a, "length"),
			detail::argument_or<const R*>(a, "initial_value", nullptr)));
	}
};
template <unsigned I, unsigned R, unsigned D> using kernel_t = detail::described_kernel_t<traits<I, R, D>>;
} // namespace scan
} // namespace reduction

} // namespace kernels
} // namespace cuda

#endif // GIDDY_COMPRESSION_WRAPPERS_HPP_
