"""libgiddy_b200 -- a Blackwell-native (sm_100a) implementation of the decompression path of
eyalroz/libgiddy, behind Giddy's own kernel-wrapper / factory interface.

The product is the shared library libgiddy_b200/lib/libgiddy_b200.so (C ABI: include/giddy_b200.h,
C++ API: include/giddy/).  This Python package is a thin ctypes binding of the library's registry
API, used by the test-suite and the benchmark; there is no Python or CPU fallback for any kernel.
"""
from ._native import (  # noqa: F401
    Library, Kernel, Scan, ScanChunk, LaunchConfig, make_launch_config, DeviceProperties, FunctionAttributes,
    b200_device_properties, nominal_function_attributes,
    GiddyError, InvalidArgument, OutOfRange, BadAnyCast, LogicError, CudaError, Unsupported,
    u8, u16, u32, u64, i8, i16, i32, i64, size_t, unsigned, serialization_factor, boolean,
    uint_of_size, size_type_by_index_size, coefficients, ptr, cptr,
)
from .schemes import default_library  # noqa: F401


def produce_subclass(key):
    """cuda::registered::kernel_t::produceSubclass(key) on the product library"""
    return default_library().produce_subclass(key)


def registered_keys():
    return default_library().registered_keys()
