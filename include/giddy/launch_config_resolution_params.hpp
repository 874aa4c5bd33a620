// giddy/launch_config_resolution_params.hpp -- the declarative half of the uniform
// launch-configuration mechanism: each kernel describes how its grid must be built, and one
// resolver (resolve_launch_configuration.hpp) turns that description into (grid, block, shared memory).
//
// Mirrors cuda::kernels::launch_config_resolution_params_t of the reference,
// src/kernels/launch_config_resolution_params.h:12-134 -- same field names and meanings, so code
// which fills in or inspects these parameters ("mode i") carries over.
#pragma once
#ifndef GIDDY_LAUNCH_CONFIG_RESOLUTION_PARAMS_HPP_
#define GIDDY_LAUNCH_CONFIG_RESOLUTION_PARAMS_HPP_

#include "types.hpp"

namespace cuda {
namespace kernels {

class launch_config_resolution_params_t {
public:
	/// Which part of the grid handles one "length unit" of the input
	enum resolution_t {
		thread, ///< threads may be grouped into blocks (almost) arbitrarily
		warp,   ///< blocks must be made of full warps; one unit per warp
		block   ///< the kernel has strict expectations of a block; one unit per block
	};

	/// Whether a thread/warp/block processes more than one length unit, and how
	enum serialization_option_t {
		none,                     ///< one length unit per thread/warp/block
		fixed_factor,             ///< a factor hard-coded in the kernel
		runtime_specified_factor, ///< a factor the kernel takes at run time
		auto_maximized,           ///< grid-stride until the input ends: use just enough blocks to keep the GPU busy
		keep_gpu_busy = auto_maximized
	};

	device::properties_t            device_properties;
	device_function::attributes_t   kernel_function_attributes;
	::std::size_t                   length { 0 }; ///< in units of the resolution
	resolution_t                    grid_construction_resolution { thread };
	serialization_option_t          serialization_option { none };
	serialization_factor_t          default_serialization_factor { 1 };

	struct {
		grid_block_dimension_t            threads_in_block { 1 };
		optional<grid_block_dimension_t>  warps_in_block   { nullopt };
		optional<grid_dimension_t>        blocks_in_grid   { nullopt };
	} quanta; ///< the number of threads in a block must be a multiple of these

	struct {
		optional<grid_block_dimension_t>  threads_in_block { nullopt };
		optional<grid_block_dimension_t>  warps_in_block   { nullopt };
	} upper_limits;

	bool                            must_have_power_of_2_threads_per_block { false };

	struct {
		optional<shared_memory_size_t>  per_length_unit { nullopt }; ///< per thread of the block
		optional<shared_memory_size_t>  per_block       { nullopt };
	} dynamic_shared_memory_requirement;

	struct { ///< meaningful only with block resolution
		optional<grid_block_dimension_t>  fixed_threads_per_block { nullopt };
		optional<grid_block_dimension_t>  max_threads_per_block   { nullopt };
	} block_resolution_constraints;

	shared_memory_size_t            available_dynamic_shared_memory_per_block;

	/// With auto_maximized serialization the grid is capped at
	/// SMs x warp schedulers per SM x this factor / warps per block.
	unsigned                        keep_gpu_busy_factor { 50 };

	bool use_dynamic_shared_memory() const
	{
		return dynamic_shared_memory_requirement.per_block || dynamic_shared_memory_requirement.per_length_unit;
	}

protected:
	launch_config_resolution_params_t(
		device::properties_t            device_properties_,
		device_function::attributes_t   kernel_function_attributes_,
		optional<shared_memory_size_t>  shared_memory_size_limit = nullopt) :
		device_properties(device_properties_), kernel_function_attributes(kernel_function_attributes_)
	{
		available_dynamic_shared_memory_per_block =
			device_function::maximum_dynamic_shared_memory_per_block(kernel_function_attributes, device_properties);
		if (shared_memory_size_limit && shared_memory_size_limit.value() < available_dynamic_shared_memory_per_block) {
			available_dynamic_shared_memory_per_block = shared_memory_size_limit.value();
		}
	}

	launch_config_resolution_params_t(
		device::properties_t            device_properties_,
		device_function_t               device_function,
		optional<shared_memory_size_t>  shared_memory_size_limit = nullopt) :
		launch_config_resolution_params_t(device_properties_, device_function.attributes(), shared_memory_size_limit) { }
};

} // namespace kernels
} // namespace cuda

#endif // GIDDY_LAUNCH_CONFIG_RESOLUTION_PARAMS_HPP_
