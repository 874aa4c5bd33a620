// giddy/resolve_launch_configuration.hpp -- the resolver half of the uniform launch-configuration
// mechanism.  Mirrors cuda::kernels::resolve_launch_configuration of the reference,
// src/kernels/resolve_launch_configuration.h:22-25.
#pragma once
#ifndef GIDDY_RESOLVE_LAUNCH_CONFIGURATION_HPP_
#define GIDDY_RESOLVE_LAUNCH_CONFIGURATION_HPP_

#include "launch_config_resolution_params.hpp"
#include "kernel_wrapper.hpp" // launch_configuration_limits_t

namespace cuda {
namespace kernels {

/// @param params  what the kernel needs of its grid
/// @param limits  one-off limits on the result
/// @param serialization_factor  for kernels which take one at run time
/// @throws std::invalid_argument when the constraints cannot be met, std::logic_error when a
/// serialization factor is given to a kernel which does not admit one
GIDDY_API launch_configuration_t resolve_launch_configuration(
	const launch_config_resolution_params_t&  params,
	const launch_configuration_limits_t&      limits = { nullopt, nullopt, nullopt },
	optional<serialization_factor_t>          serialization_factor = nullopt);

} // namespace kernels
} // namespace cuda

#endif // GIDDY_RESOLVE_LAUNCH_CONFIGURATION_HPP_
