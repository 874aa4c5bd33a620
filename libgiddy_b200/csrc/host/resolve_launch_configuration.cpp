// The launch-configuration resolver: declarative per-kernel parameters -> (grid, block, shared memory).
//
// Same algorithm, step for step, as the reference's src/kernels/resolve_launch_configuration.cpp:114-409
// (SURVEY.md Appendix F), written afresh:
//   1. effective length = length / serialization factor                       (reference :345-377)
//   2. threads per block = the largest count every constraint admits, rounded
//      down to the block-length quantum                                       (reference :136-260)
//   3. blocks = ceil(effective length / units per block), capped for
//      auto-maximized kernels at what keeps the GPU busy                      (reference :262-332)
//   4. dynamic shared memory = per block + threads x per length unit          (reference :334-343)
// One deliberate difference: the cap on threads by the number of length units is
// round_up(units x threads-per-unit, quantum); the reference passes these two arguments swapped
// (:171-173), which only matters for inputs smaller than one block.
#include "resolve_launch_configuration.hpp"

#include <algorithm>
#include <limits>
#include <numeric>
#include <string>

namespace cuda {
namespace kernels {

namespace {

using params_t     = launch_config_resolution_params_t;
using limits_t     = launch_configuration_limits_t;
using resolution_t = params_t::resolution_t;

constexpr ::std::size_t unconstrained = ::std::numeric_limits<::std::size_t>::max();

::std::size_t round_down_to_power_of_2(::std::size_t x)
{
	if (x == 0) { return 0; }
	::std::size_t p = 1;
	while (p <= x / 2) { p *= 2; }
	return p;
}

::std::size_t effective_length(const params_t& params, optional<serialization_factor_t> requested_factor)
{
	::std::size_t factor;
	switch (params.serialization_option) {
	case params_t::auto_maximized:
		if (requested_factor) {
			throw ::std::logic_error("kernel auto-serializes, making it meaningless to apply a serialization factor");
		}
		return params.length;
	case params_t::none:
		if (requested_factor.value_or(1) != 1) { throw ::std::logic_error("kernel does not allow for serialization"); }
		return params.length;
	case params_t::fixed_factor:
		factor = requested_factor.value_or(params.default_serialization_factor);
		if (factor != params.default_serialization_factor) {
			throw ::std::logic_error("kernel cannot admit serialization at a factor other than the fixed default");
		}
		break;
	case params_t::runtime_specified_factor:
		factor = requested_factor.value_or(params.default_serialization_factor);
		break;
	default:
		throw ::std::invalid_argument("Unsupported serialization option");
	}
	if (factor == 0) { throw ::std::invalid_argument("A serialization factor must be positive"); }
	return params.length == 0 ? 0 : 1 + (params.length - 1) / factor;
}

::std::size_t block_length_quantum(const params_t& params)
{
	if (params.grid_construction_resolution == resolution_t::block && params.block_resolution_constraints.fixed_threads_per_block) {
		return params.block_resolution_constraints.fixed_threads_per_block.value();
	}
	::std::size_t quantum = params.quanta.threads_in_block;
	if (params.quanta.warps_in_block) { quantum = ::std::lcm<::std::size_t>(quantum, params.quanta.warps_in_block.value() * warp_size); }
	if (params.grid_construction_resolution == resolution_t::warp) { quantum = ::std::lcm<::std::size_t>(quantum, (::std::size_t) warp_size); }
	return quantum;
}

::std::size_t max_threads_per_block(const params_t& params, const limits_t& limits, ::std::size_t eff_length)
{
	const ::std::size_t by_device = (::std::size_t) params.device_properties.maxThreadsDim[0];

	::std::size_t threads_per_unit;
	switch (params.grid_construction_resolution) {
	case resolution_t::thread: threads_per_unit = 1; break;
	case resolution_t::warp:   threads_per_unit = warp_size; break;
	case resolution_t::block:  threads_per_unit = params.block_resolution_constraints.max_threads_per_block.value_or(by_device); break;
	default: throw ::std::invalid_argument("Unsupported grid construction resolution");
	}
	const ::std::size_t quantum = block_length_quantum(params);
	::std::size_t by_num_length_units = unconstrained;
	if (quantum > 0 && eff_length < unconstrained / threads_per_unit) {
		by_num_length_units = util::round_up<::std::size_t>(eff_length * threads_per_unit, quantum);
	}

	::std::size_t by_dynamic_shared_memory = unconstrained;
	if (params.use_dynamic_shared_memory()) {
		const ::std::size_t per_block = params.dynamic_shared_memory_requirement.per_block.value_or(0);
		const ::std::size_t per_unit = params.dynamic_shared_memory_requirement.per_length_unit.value_or(0);
		if (params.available_dynamic_shared_memory_per_block < per_block) { by_dynamic_shared_memory = 0; }
		else if (per_unit != 0) { by_dynamic_shared_memory = (params.available_dynamic_shared_memory_per_block - per_block) / per_unit; }
	}

	::std::size_t by_device_function = by_device;
	switch (params.grid_construction_resolution) {
	case resolution_t::thread: by_device_function = (::std::size_t) params.kernel_function_attributes.maxThreadsPerBlock; break;
	case resolution_t::warp:   by_device_function = util::round_down<::std::size_t>((::std::size_t) params.kernel_function_attributes.maxThreadsPerBlock, warp_size); break;
	default: break; // block-resolution kernels: the reference does not consult the function attribute
	}

	::std::size_t result = by_device;
	result = ::std::min(result, by_dynamic_shared_memory);
	result = ::std::min(result, by_device_function);
	result = ::std::min(result, by_num_length_units);
	result = ::std::min(result, (::std::size_t) limits.block_size.value_or(::std::numeric_limits<grid_block_dimension_t>::max()));
	if (params.upper_limits.threads_in_block) { result = ::std::min(result, (::std::size_t) params.upper_limits.threads_in_block.value()); }
	if (params.upper_limits.warps_in_block)   { result = ::std::min(result, (::std::size_t) params.upper_limits.warps_in_block.value() * warp_size); }
	if (params.must_have_power_of_2_threads_per_block) { result = round_down_to_power_of_2(result); }
	return result;
}

grid_block_dimension_t block_length(const params_t& params, ::std::size_t max_threads)
{
	if (params.grid_construction_resolution == resolution_t::block && params.block_resolution_constraints.fixed_threads_per_block) {
		auto fixed = params.block_resolution_constraints.fixed_threads_per_block.value();
		if (fixed > max_threads) {
			throw ::std::invalid_argument("Kernel requires " + ::std::to_string(fixed) + " threads per block, while the different "
				"constraints on the number of threads per block only allow " + ::std::to_string(max_threads) + ".");
		}
		return fixed;
	}
	const ::std::size_t quantum = block_length_quantum(params);
	const ::std::size_t threads = (quantum == 0) ? 0 : util::round_down(max_threads, quantum);
	if (threads == 0) {
		throw ::std::invalid_argument("The combined constraints on the number of threads per block do not allow any threads in a block");
	}
	return (grid_block_dimension_t) threads;
}

::std::size_t blocks_to_keep_gpu_busy(const params_t& params, grid_block_dimension_t threads_per_block)
{
	const ::std::size_t warps_per_block = util::div_rounding_up<::std::size_t>(threads_per_block, warp_size);
	return (::std::size_t) params.device_properties.multiProcessorCount *
		params.device_properties.compute_capability().max_warp_schedulings_per_processor_cycle() *
		params.keep_gpu_busy_factor / warps_per_block;
}

::std::size_t num_blocks(const params_t& params, ::std::size_t eff_length, grid_block_dimension_t threads_per_block, optional<grid_dimension_t> max_blocks)
{
	::std::size_t units_per_block;
	switch (params.grid_construction_resolution) {
	case resolution_t::thread: units_per_block = threads_per_block; break;
	case resolution_t::warp:   units_per_block = threads_per_block / warp_size; break;
	case resolution_t::block:  units_per_block = 1; break;
	default: throw ::std::invalid_argument("Unsupported grid construction resolution");
	}
	const ::std::size_t covering = util::div_rounding_up(eff_length, units_per_block);
	const ::std::size_t limit = max_blocks ? (::std::size_t) max_blocks.value() : covering;
	if (params.serialization_option == params_t::auto_maximized) {
		return ::std::min(::std::min(covering, blocks_to_keep_gpu_busy(params, threads_per_block)), limit);
	}
	if (covering > limit) {
		throw ::std::invalid_argument("Required grid length exceeds specified limit: " + ::std::to_string(limit) + " > " + ::std::to_string(covering));
	}
	return covering;
}

} // namespace

launch_configuration_t resolve_launch_configuration(
	const launch_config_resolution_params_t&  params,
	const launch_configuration_limits_t&      limits,
	optional<serialization_factor_t>          serialization_factor)
{
	const ::std::size_t eff_length = effective_length(params, serialization_factor);
	const ::std::size_t max_threads = max_threads_per_block(params, limits, eff_length);
	const grid_block_dimension_t threads = block_length(params, max_threads);
	const ::std::size_t blocks = num_blocks(params, eff_length, threads, limits.grid_size);
	if (blocks > (::std::size_t) ::std::numeric_limits<int>::max()) {
		throw ::std::invalid_argument("The input requires more blocks than a grid can have");
	}
	const ::std::size_t shared_memory = (::std::size_t) params.dynamic_shared_memory_requirement.per_block.value_or(0) +
		(::std::size_t) threads * params.dynamic_shared_memory_requirement.per_length_unit.value_or(0);
	if (limits.dynamic_shared_memory && shared_memory > limits.dynamic_shared_memory.value()) {
		throw ::std::invalid_argument("The kernel needs more dynamic shared memory than the specified limit allows");
	}
	return make_launch_config(dimensions_t((unsigned) blocks), dimensions_t(threads), (shared_memory_size_t) shared_memory);
}

} // namespace kernels
} // namespace cuda
