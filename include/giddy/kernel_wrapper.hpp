// giddy/kernel_wrapper.hpp -- the abstract kernel wrapper: a kernel together with the code which
// works out its launch configuration, with type-erased ("virtual") argument passing.
//
// Mirrors cuda::kernel_t and cuda::launch_configuration_limits_t of the reference,
// src/cuda/kernel_wrapper.cuh:20-128, and the free functions of namespace cuda::kernel, :130-215.
#pragma once
#ifndef GIDDY_KERNEL_WRAPPER_HPP_
#define GIDDY_KERNEL_WRAPPER_HPP_

#include "types.hpp"

#include <string>
#include <unordered_map>

namespace cuda {

/// One-off limits imposed on a launch configuration ("used for testing, essentially")
struct launch_configuration_limits_t {
	optional<grid_block_dimension_t> block_size;
	optional<grid_dimension_t>       grid_size;
	optional<shared_memory_size_t>   dynamic_shared_memory;

	static launch_configuration_limits_t limit_block_size(optional<grid_block_dimension_t> limit) { return { limit, nullopt, nullopt }; }
	bool none() const { return !block_size && !grid_size && !dynamic_shared_memory; }
};

/// No data members, just a vtable: wrapper objects are stateless.
struct kernel_t {
	/// name -> value; values must have *exactly* the type the wrapper expects (any_cast)
	using arguments_type = ::std::unordered_map<::std::string, any>;

	/// Enqueue the kernel on a stream of the current device.  Asynchronous and, as in the reference,
	/// unchecked with respect to execution errors.
	virtual void enqueue_launch(
		stream::id_t                    stream_id,
		const launch_configuration_t&   launch_config,
		arguments_type                  args) const = 0;

	/// Resolve a launch configuration from meta-data about the inputs; fetches the kernel's
	/// attributes itself.
	virtual launch_configuration_t resolve_launch_configuration(
		device::properties_t            device_properties,
		arguments_type                  extra_arguments,
		launch_configuration_limits_t   limits = { nullopt, nullopt, nullopt }) const
	{
		return resolve_launch_configuration(device_properties, get_device_function().attributes(), extra_arguments, limits);
	}

	virtual launch_configuration_t resolve_launch_configuration(
		device::properties_t            device_properties,
		device_function::attributes_t   kernel_function_attributes,
		arguments_type                  extra_arguments,
		launch_configuration_limits_t   limits = { nullopt, nullopt, nullopt }) const = 0;

	virtual const device_function_t get_device_function() const = 0;

	virtual ~kernel_t() { }
};

namespace kernel {

/// "Mode ii": launch a wrapper's kernel with a caller-assembled parameter pack.
template <typename... KernelParameters>
inline void enqueue_launch(const kernel_t& kernel, stream::id_t stream_id, const launch_configuration_t& launch_config, KernelParameters... parameters)
{
	cuda::enqueue_launch(kernel.get_device_function(), stream_id, launch_config, parameters...);
}

template <typename... KernelParameters>
inline void enqueue_launch(const kernel_t& kernel, const launch_configuration_t& launch_config, KernelParameters... parameters)
{
	cuda::enqueue_launch(kernel.get_device_function(), stream::default_stream_id, launch_config, parameters...);
}

inline launch_configuration_t resolve_launch_configuration(
	const kernel_t& kernel, device::properties_t device_properties, kernel_t::arguments_type extra_arguments, launch_configuration_limits_t limits = { })
{
	return kernel.resolve_launch_configuration(device_properties, kernel.get_device_function().attributes(), extra_arguments, limits);
}

inline launch_configuration_t resolve_launch_configuration(
	const kernel_t& kernel, device::id_t device_id, kernel_t::arguments_type extra_arguments, launch_configuration_limits_t limits = { })
{
	return resolve_launch_configuration(kernel, device::get_properties(device_id), extra_arguments, limits);
}

} // namespace kernel
} // namespace cuda

#endif // GIDDY_KERNEL_WRAPPER_HPP_
