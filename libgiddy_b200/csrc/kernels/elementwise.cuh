// Position-aligned ("one input element -> one output element at the same index") decoders:
// Frame of Reference, Model, Discard-Zero-Bytes fixed width, Dictionary.
//
// Shared skeleton: a thread produces whole 16-byte output vectors (E = 16 / sizeof(Uncompressed)
// elements), reading the E matching compressed elements with one E*sizeof(Compressed)-byte
// access.  Consecutive lanes own consecutive vectors, so every warp-level store covers 512
// contiguous bytes and every warp-level load 32*E*C contiguous bytes.  Each thread keeps
// `unroll` independent loads in flight before it starts computing.  A CTA owns one tile of
// blockDim.x * unroll vectors per grid-stride iteration.  Misaligned buffers and the < E
// trailing elements take an element-wise path.
#pragma once
#ifndef GIDDY_B200_KERNELS_ELEMENTWISE_CUH_
#define GIDDY_B200_KERNELS_ELEMENTWISE_CUH_

#include "common.cuh"

namespace giddy {
namespace kernels {

using namespace giddy::device;

namespace detail {

constexpr unsigned elementwise_unroll = GIDDY_ELEMENTWISE_UNROLL;

// position / segment_length and position % segment_length, avoiding wide or slow divisions
// where the operands allow.
template <typename Size>
struct segment_locator {
	Size     segment_length;
	unsigned log2_segment_length; // valid only when is_power_of_2
	bool     is_power_of_2;

	__device__ __forceinline__ explicit segment_locator(Size L) : segment_length(L)
	{
		is_power_of_2 = (L & (L - 1)) == 0;
		log2_segment_length = (sizeof(Size) > 4) ? (63 - __clzll((long long) L)) : (31 - __clz((int) L));
	}

	__device__ __forceinline__ void locate(Size position, Size& segment_index, Size& position_in_segment) const
	{
		if (is_power_of_2) {
			segment_index = position >> log2_segment_length;
			position_in_segment = position & (segment_length - 1);
		}
		else if (sizeof(Size) > 4 && ((position | segment_length) >> 32) == 0) {
			uint32_t q = (uint32_t) position / (uint32_t) segment_length;
			segment_index = q;
			position_in_segment = (uint32_t) position - q * (uint32_t) segment_length;
		}
		else {
			segment_index = position / segment_length;
			position_in_segment = position - segment_index * segment_length;
		}
	}
};

// Evaluate a parametric model; see oracle/giddy_oracle.cpp eval_model and
// the reference's src/cuda/model_functors.hpp:87-90,108-111.
template <int Model, unsigned ModelIndexSize, typename Coefficient, unsigned ResultSize, typename Position>
__device__ __forceinline__ uint_t<ResultSize> evaluate_model(const Coefficient* __restrict__ c, Position x)
{
	if (Model == model_constant) { return (uint_t<ResultSize>) c[0]; }
	if (ResultSize > 4) { // x is first narrowed to the model's index type
		uint64_t xi = (ModelIndexSize > 4) ? (uint64_t) x : (uint64_t)(uint32_t) x;
		return (uint_t<ResultSize>)((uint64_t)(int64_t) c[0] + (uint64_t)(int64_t) c[1] * xi);
	}
	// the result's low 32 bits depend only on the operands' low 32 bits
	return (uint_t<ResultSize>)((uint32_t) c[0] + (uint32_t) c[1] * (uint32_t) x);
}

} // namespace detail

// -------------------------------------------------------------------------------------------------
// Frame of Reference.  Replaces cuda::kernels::decompression::frame_of_reference::decompress,
// reference src/kernels/decompression/frame_of_reference.cuh:35-80 (same parameter list; the
// last parameter is a launch-shape hint of the reference's and is ignored here).
//   decompressed[i] = (U)( (U) model_s(i - s*L) + compressed_input[i] ),  s = i / L
// Coefficients are int32, `Dim` per segment (1: constant, 2: linear).
// -------------------------------------------------------------------------------------------------
template <unsigned IndexSize, unsigned U, unsigned C, int Model, unsigned ModelIndexSize>
__global__ void frame_of_reference_decompress(
	uint_t<U>*          __restrict__  decompressed,
	const int_t<C>*     __restrict__  compressed_input,
	const int32_t*      __restrict__  segment_model_coefficients,
	size_type<IndexSize>              length,
	size_type<IndexSize>              segment_length,
	uint_t<IndexSize>                 /* num_segments_to_decompress_per_block */)
{
	using size_t_ = size_type<IndexSize>;
	constexpr unsigned E = 16 / U;
	constexpr unsigned in_bytes = E * C;
	constexpr unsigned dim = (Model == model_linear) ? 2 : 1;
	constexpr unsigned unroll = detail::elementwise_unroll;

	const detail::segment_locator<size_t_> locator(segment_length);

	auto decode_one = [&](size_t_ position) {
		size_t_ s, x;
		locator.locate(position, s, x);
		const int32_t* c = segment_model_coefficients + (size_t) s * dim;
		int32_t coefficients[dim];
		#pragma unroll
		for (unsigned d = 0; d < dim; d++) { coefficients[d] = __ldg(c + d); }
		uint_t<U> predicted = detail::evaluate_model<Model, ModelIndexSize, int32_t, U>(coefficients, x);
		decompressed[position] = (uint_t<U>)(predicted + (uint_t<U>)(int_t<U>) compressed_input[position]);
	};

	const bool vectorizable = is_aligned_to(decompressed, 16) && is_aligned_to(compressed_input, in_bytes);
	const uint64_t global_thread = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;

	if (!vectorizable) {
		for (uint64_t p = global_thread; p < length; p += (uint64_t) gridDim.x * blockDim.x) { decode_one((size_t_) p); }
		return;
	}

	const uint64_t num_vectors = length / E;
	const uint64_t vectors_per_tile = (uint64_t) blockDim.x * unroll;

	#pragma unroll 1
	for (uint64_t tile_start = (uint64_t) blockIdx.x * vectors_per_tile; tile_start < num_vectors;
	     tile_start += (uint64_t) gridDim.x * vectors_per_tile)
	{
		byte_pack<in_bytes> in[unroll];
		#pragma unroll
		for (unsigned k = 0; k < unroll; k++) {
			uint64_t v = tile_start + (uint64_t) k * blockDim.x + threadIdx.x;
			if (v < num_vectors) { in[k] = load_pack<in_bytes>(compressed_input + v * E); }
		}
		#pragma unroll
		for (unsigned k = 0; k < unroll; k++) {
			uint64_t v = tile_start + (uint64_t) k * blockDim.x + threadIdx.x;
			if (v >= num_vectors) { continue; }
			size_t_ s, x;
			locator.locate((size_t_)(v * E), s, x);
			int32_t coefficients[dim];
			#pragma unroll
			for (unsigned d = 0; d < dim; d++) { coefficients[d] = __ldg(segment_model_coefficients + (size_t) s * dim + d); }
			byte_pack<16> out;
			#pragma unroll
			for (unsigned j = 0; j < E; j++) {
				uint_t<U> predicted = detail::evaluate_model<Model, ModelIndexSize, int32_t, U>(coefficients, x);
				uint_t<U> offset = (uint_t<U>)(int_t<U>)(int_t<C>) in[k].template get<C>(j);
				out.template set<U>(j, (uint_t<U>)(predicted + offset));
				x++;
				if (x == segment_length && j + 1 < E) { // the vector straddles a segment boundary (L % E != 0)
					x = 0; s++;
					#pragma unroll
					for (unsigned d = 0; d < dim; d++) { coefficients[d] = __ldg(segment_model_coefficients + (size_t) s * dim + d); }
				}
			}
			store_pack_16(decompressed + v * E, out);
		}
	}
	// the < E trailing elements
	uint64_t tail_position = num_vectors * E + global_thread;
	if (tail_position < length) { decode_one((size_t_) tail_position); }
}

// -------------------------------------------------------------------------------------------------
// Model.  Replaces cuda::kernels::decompression::model::decompress,
// reference src/kernels/decompression/model.cuh:38-53:  decompressed[i] = (U) model(i).
// One model for the whole column; coefficients (of the uncompressed type) are passed by value.
// -------------------------------------------------------------------------------------------------
template <unsigned IndexSize, unsigned U, int Model>
__global__ void model_decompress(
	uint_t<U>* __restrict__                                   decompressed,
	pod_array<int_t<U>, (Model == model_linear ? 2u : 1u)>    model_coefficients,
	size_type<IndexSize>                                      length)
{
	constexpr unsigned E = 16 / U;
	constexpr unsigned unroll = detail::elementwise_unroll;
	const uint64_t global_thread = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;

	if (!is_aligned_to(decompressed, 16)) {
		for (uint64_t p = global_thread; p < length; p += (uint64_t) gridDim.x * blockDim.x) {
			decompressed[p] = detail::evaluate_model<Model, IndexSize, int_t<U>, U>(model_coefficients.v, p);
		}
		return;
	}
	const uint64_t num_vectors = length / E;
	const uint64_t vectors_per_tile = (uint64_t) blockDim.x * unroll;
	#pragma unroll 1
	for (uint64_t tile_start = (uint64_t) blockIdx.x * vectors_per_tile; tile_start < num_vectors;
	     tile_start += (uint64_t) gridDim.x * vectors_per_tile)
	{
		#pragma unroll
		for (unsigned k = 0; k < unroll; k++) {
			uint64_t v = tile_start + (uint64_t) k * blockDim.x + threadIdx.x;
			if (v >= num_vectors) { continue; }
			byte_pack<16> out;
			#pragma unroll
			for (unsigned j = 0; j < E; j++) {
				out.template set<U>(j, detail::evaluate_model<Model, IndexSize, int_t<U>, U>(model_coefficients.v, v * E + j));
			}
			store_pack_16(decompressed + v * E, out);
		}
	}
	uint64_t tail_position = num_vectors * E + global_thread;
	if (tail_position < length) {
		decompressed[tail_position] = detail::evaluate_model<Model, IndexSize, int_t<U>, U>(model_coefficients.v, tail_position);
	}
}

// -------------------------------------------------------------------------------------------------
// Model with intervals.  Replaces cuda::kernels::decompression::model::decompress_with_intervals, reference
// src/kernels/decompression/model.cuh:55-98 (a kernel the reference has but never wraps or registers):
//   decompressed[i] = (U) model_s(i - s * modeling_period),  s = i / modeling_period
// -- Frame of Reference without the offsets.  Coefficients are of the uncompressed type, `dim` per interval.
// -------------------------------------------------------------------------------------------------
template <unsigned IndexSize, unsigned U, int Model>
__global__ void model_decompress_with_intervals(
	uint_t<U>*        __restrict__  decompressed,
	const int_t<U>*   __restrict__  interval_model_coefficients,
	size_type<IndexSize>            modeling_period,
	size_type<IndexSize>            length)
{
	using size_t_ = size_type<IndexSize>;
	constexpr unsigned E = 16 / U;
	constexpr unsigned dim = (Model == model_linear) ? 2 : 1;
	constexpr unsigned unroll = detail::elementwise_unroll;
	const detail::segment_locator<size_t_> locator(modeling_period);
	const uint64_t global_thread = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	auto value_at = [&](size_t_ position) {
		size_t_ s, x;
		locator.locate(position, s, x);
		int_t<U> c[dim];
		#pragma unroll
		for (unsigned d = 0; d < dim; d++) { c[d] = __ldg(interval_model_coefficients + (size_t) s * dim + d); }
		return detail::evaluate_model<Model, IndexSize, int_t<U>, U>(c, x);
	};
	if (!is_aligned_to(decompressed, 16)) {
		for (uint64_t p = global_thread; p < length; p += (uint64_t) gridDim.x * blockDim.x) { decompressed[p] = value_at((size_t_) p); }
		return;
	}
	const uint64_t num_vectors = length / E;
	const uint64_t vectors_per_tile = (uint64_t) blockDim.x * unroll;
	#pragma unroll 1
	for (uint64_t tile_start = (uint64_t) blockIdx.x * vectors_per_tile; tile_start < num_vectors; tile_start += (uint64_t) gridDim.x * vectors_per_tile) {
		#pragma unroll
		for (unsigned k = 0; k < unroll; k++) {
			const uint64_t v = tile_start + (uint64_t) k * blockDim.x + threadIdx.x;
			if (v >= num_vectors) { continue; }
			size_t_ s, x;
			locator.locate((size_t_)(v * E), s, x);
			int_t<U> c[dim];
			#pragma unroll
			for (unsigned d = 0; d < dim; d++) { c[d] = __ldg(interval_model_coefficients + (size_t) s * dim + d); }
			byte_pack<16> out;
			#pragma unroll
			for (unsigned j = 0; j < E; j++) {
				out.template set<U>(j, detail::evaluate_model<Model, IndexSize, int_t<U>, U>(c, x));
				x++;
				if (x == modeling_period && j + 1 < E) { // the vector straddles an interval boundary
					x = 0; s++;
					#pragma unroll
					for (unsigned d = 0; d < dim; d++) { c[d] = __ldg(interval_model_coefficients + (size_t) s * dim + d); }
				}
			}
			store_pack_16(decompressed + v * E, out);
		}
	}
	const uint64_t tail_position = num_vectors * E + global_thread;
	if (tail_position < length) { decompressed[tail_position] = value_at((size_t_) tail_position); }
}

// -------------------------------------------------------------------------------------------------
// Discard zero bytes, fixed width.  Replaces
// cuda::kernels::decompression::discard_zero_bytes::fixed_width::decompress,
// reference src/kernels/decompression/discard_zero_bytes_fixed.cuh:108-124 (little-endian):
//   decompressed[i] = zero_extend(compressed_input[i])
// -------------------------------------------------------------------------------------------------
template <unsigned IndexSize, unsigned U, unsigned C>
__global__ void discard_zero_bytes_fixed_decompress(
	uint_t<U>*        __restrict__  decompressed,
	const uint_t<C>*  __restrict__  compressed_input,
	size_type<IndexSize>            length)
{
	constexpr unsigned E = 16 / U;
	constexpr unsigned in_bytes = E * C;
	constexpr unsigned unroll = detail::elementwise_unroll;
	const uint64_t global_thread = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;

	if (!(is_aligned_to(decompressed, 16) && is_aligned_to(compressed_input, in_bytes))) {
		for (uint64_t p = global_thread; p < length; p += (uint64_t) gridDim.x * blockDim.x) {
			decompressed[p] = (uint_t<U>) compressed_input[p];
		}
		return;
	}
	const uint64_t num_vectors = length / E;
	const uint64_t vectors_per_tile = (uint64_t) blockDim.x * unroll;
	#pragma unroll 1
	for (uint64_t tile_start = (uint64_t) blockIdx.x * vectors_per_tile; tile_start < num_vectors;
	     tile_start += (uint64_t) gridDim.x * vectors_per_tile)
	{
		byte_pack<in_bytes> in[unroll];
		#pragma unroll
		for (unsigned k = 0; k < unroll; k++) {
			uint64_t v = tile_start + (uint64_t) k * blockDim.x + threadIdx.x;
			if (v < num_vectors) { in[k] = load_pack<in_bytes>(compressed_input + v * E); }
		}
		#pragma unroll
		for (unsigned k = 0; k < unroll; k++) {
			uint64_t v = tile_start + (uint64_t) k * blockDim.x + threadIdx.x;
			if (v >= num_vectors) { continue; }
			byte_pack<16> out;
			#pragma unroll
			for (unsigned j = 0; j < E; j++) { out.template set<U>(j, (uint_t<U>) in[k].template get<C>(j)); }
			store_pack_16(decompressed + v * E, out);
		}
	}
	uint64_t tail_position = num_vectors * E + global_thread;
	if (tail_position < length) { decompressed[tail_position] = (uint_t<U>) compressed_input[tail_position]; }
}

// -------------------------------------------------------------------------------------------------
// Dictionary.  Replaces cuda::kernels::decompression::dictionary::decompress,
// reference src/kernels/decompression/dictionary.cuh:61-75 -> gather.cuh:16-146 (same parameters):
//   decompressed[i] = dictionary_entries[compressed_input[i]]
// When cache_dictionary_in_shared_memory is set the launch must provide
// num_dictionary_entries * U bytes of dynamic shared memory; the dictionary is then read from
// HBM exactly once per CTA and every lookup is a shared-memory access.  Otherwise lookups go
// through the read-only L1/L2 path.
// -------------------------------------------------------------------------------------------------
template <unsigned IndexSize, unsigned U, unsigned DictionaryIndexSize>
__global__ void dictionary_decompress(
	uint_t<U>*                          __restrict__  decompressed,
	const uint_t<DictionaryIndexSize>*  __restrict__  compressed_input,
	const uint_t<U>*                    __restrict__  dictionary_entries,
	size_type<IndexSize>                              length,
	size_type<DictionaryIndexSize>                    num_dictionary_entries,
	bool                                              cache_dictionary_in_shared_memory,
	unsigned short                                    /* serialization_factor: unused, as in the reference */)
{
	constexpr unsigned DI = DictionaryIndexSize;
	constexpr unsigned E = 16 / U;
	constexpr unsigned in_bytes = E * DI;
	constexpr unsigned unroll = detail::elementwise_unroll;
	const uint64_t global_thread = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;

	// As many of the dictionary's leading entries as the launch's dynamic shared memory holds are served from there
	// (all of them, normally; a dictionary which does not fit keeps its head -- where a frequency-ordered dictionary
	// has its hot entries -- in shared memory and its tail behind the read-only L1/L2 path).
	const uint_t<U>* const cached = reinterpret_cast<const uint_t<U>*>(dynamic_shared_memory);
	uint64_t num_cached = 0;
	if (cache_dictionary_in_shared_memory) {
		num_cached = minimum<uint64_t>(num_dictionary_entries, dynamic_shared_memory_size() / U);
		auto destination = reinterpret_cast<uint_t<U>*>(dynamic_shared_memory);
		for (uint64_t i = threadIdx.x; i < num_cached; i += blockDim.x) { destination[i] = __ldg(dictionary_entries + i); }
		__syncthreads();
	}

	auto decode = [&](auto lookup) {
		if (!(is_aligned_to(decompressed, 16) && is_aligned_to(compressed_input, in_bytes))) {
			for (uint64_t p = global_thread; p < length; p += (uint64_t) gridDim.x * blockDim.x) {
				decompressed[p] = lookup(compressed_input[p]);
			}
			return;
		}
		const uint64_t num_vectors = length / E;
		const uint64_t vectors_per_tile = (uint64_t) blockDim.x * unroll;
		#pragma unroll 1
		for (uint64_t tile_start = (uint64_t) blockIdx.x * vectors_per_tile; tile_start < num_vectors;
		     tile_start += (uint64_t) gridDim.x * vectors_per_tile)
		{
			byte_pack<in_bytes> in[unroll];
			#pragma unroll
			for (unsigned k = 0; k < unroll; k++) {
				uint64_t v = tile_start + (uint64_t) k * blockDim.x + threadIdx.x;
				if (v < num_vectors) { in[k] = load_pack<in_bytes>(compressed_input + v * E); }
			}
			#pragma unroll
			for (unsigned k = 0; k < unroll; k++) {
				uint64_t v = tile_start + (uint64_t) k * blockDim.x + threadIdx.x;
				if (v >= num_vectors) { continue; }
				byte_pack<16> out;
				#pragma unroll
				for (unsigned j = 0; j < E; j++) { out.template set<U>(j, lookup(in[k].template get<DI>(j))); }
				store_pack_16(decompressed + v * E, out);
			}
		}
		uint64_t tail_position = num_vectors * E + global_thread;
		if (tail_position < length) { decompressed[tail_position] = lookup(compressed_input[tail_position]); }
	};

	// One instance of the loop (three would cost its register allotment and with it two resident CTAs per
	// multiprocessor).  Fully cached is the common case: the comparison below is then uniformly true.
	if (num_cached >= num_dictionary_entries && num_cached != 0) { num_cached = ~(uint64_t) 0; }
	const uint32_t cached_limit = (uint32_t) minimum<uint64_t>(num_cached, 0xffffffffu); // indices below it are served from shared memory
	if (sizeof(uint_t<DI>) > 4 && num_cached > 0xffffffffull) {
		decode([&](uint_t<DI> index) -> uint_t<U> { return cached[index]; });
	}
	else {
		decode([&](uint_t<DI> index) -> uint_t<U> {
			return ((uint64_t) index < (uint64_t) cached_limit || cached_limit == 0xffffffffu) ? cached[index] : __ldg(dictionary_entries + index);
		});
	}
}

} // namespace kernels
} // namespace giddy

#endif // GIDDY_B200_KERNELS_ELEMENTWISE_CUH_
